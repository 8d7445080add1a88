#!/usr/bin/env python3
"""Extract the known-answer vectors the reference commits for the UDE hot path.

The reference (ChrisRackauckas/universal_differential_equations) has no test
suite for this path; what it *does* commit are the result files its scripts
wrote (`LotkaVolterra/results/*.jld2`, written by `scenario_1.jl:210-213`,
`scenario_2.jl`, `scenario_3.jl`, `hudson_bay.jl:231-235`).  Those files are
uncompressed little-endian JLD2 (HDF5 container); the arrays we need are stored
contiguously, so they can be sliced at fixed byte offsets without an HDF5
library (offset table: SURVEY.md Appendix B).  Every file is guarded by its
SHA-256 so a changed reference fails loudly instead of yielding garbage.

This script can only run where `/root/reference` is mounted (the build
container).  Its outputs, `tests/golden/*.npz`, are small and committed; the
tests and the GPU box only ever read those.

Usage:  python tools/make_golden.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import hashlib
import os

import numpy as np

FILES = {
    # name: (relative path, sha256 prefix, {key: (offset, count, dtype, shape, order)})
    "scenario_1": (
        "LotkaVolterra/results/Scenario_1_recovery_0.005.jld2",
        "a2dde7d861c9216a",
        {
            "t": (40495, 31, "<f8", (31,), "C"),
            # noisy data Xn, 2x31 column-major (per time point: x, y)
            "X": (42801, 62, "<f8", (2, 31), "F"),
            # Lux initial parameters (Float32 Glorot weights, zero biases)
            "W1_init": (48276, 10, "<f4", (5, 2), "F"),
            "W2_init": (48482, 25, "<f4", (5, 5), "F"),
            "W3_init": (48748, 25, "<f4", (5, 5), "F"),
            "W4_init": (49014, 10, "<f4", (2, 5), "F"),
            "theta_trained": (58663, 87, "<f8", (87,), "C"),
            "losses": (59442, 1097, "<f8", (1097,), "C"),
            # Xhat = predict(theta_trained, Xn[:,1], 0:0.05:3), 2x61 column-major
            "Xhat": (117992, 122, "<f8", (2, 61), "F"),
            # Yhat = U(Xhat, theta_trained), 2x61 column-major
            "Yhat": (119669, 122, "<f8", (2, 61), "F"),
            # OrdinaryDiffEq's own Tsit5 constant cache (serialized with the problem)
            "tsit5_consts": (129374, 57, "<f8", (57,), "C"),
            # OrdinaryDiffEq's own Vern7 constant cache
            "vern7_consts": (35964, 198, "<f8", (198,), "C"),
        },
    ),
    "scenario_2": (
        "LotkaVolterra/results/Scenario_2_recovery_0.005.jld2",
        "90468f9df21715fa",
        {
            "t": (43165, 61, "<f8", (61,), "C"),
            "X": (45711, 122, "<f8", (2, 61), "F"),
            "theta_init": (60162, 88, "<f8", (88,), "C"),
            "theta_trained": (60987, 88, "<f8", (88,), "C"),
            "losses": (61774, 2996, "<f8", (2996,), "C"),
            "Xhat": (133361, 242, "<f8", (2, 121), "F"),
            "Yhat": (136478, 242, "<f8", (2, 121), "F"),
        },
    ),
    "scenario_3": (
        "LotkaVolterra/results/Scenario_3_recovery_0.005.jld2",
        "a17101afdd4229ae",
        {
            "X": (17068, 286, "<f4", (26, 11), "F"),
            "theta_init": (31882, 81, "<f4", (81,), "C"),
            "theta_trained": (32327, 81, "<f4", (81,), "C"),
            "losses": (32717, 622, "<f4", (622,), "C"),
            # row-major 26x11 (built by concatenating rows, scenario_3.jl:194-195)
            "Xhat": (77931, 286, "<f4", (26, 11), "C"),
            "Rhat": (80432, 286, "<f4", (26, 11), "C"),
        },
    ),
    "hudson_bay": (
        "LotkaVolterra/results/Hudson_Bay_recovery.jld2",
        "f6a65a09d3034604",
        {
            "X": (629, 42, "<f4", (2, 21), "F"),
            "t": (862, 21, "<f4", (21,), "C"),
            "theta_init": (8504, 89, "<f4", (89,), "C"),
            "theta_trained": (8926, 89, "<f4", (89,), "C"),
            "losses": (9348, 460, "<f4", (460,), "C"),
            "Xhat": (53047, 82, "<f8", (2, 41), "F"),
            "tsample": (53769, 41, "<f8", (41,), "C"),
            "Yhat": (54171, 82, "<f8", (2, 41), "F"),
        },
    ),
}


def extract(ref_root, out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name, (rel, sha_prefix, fields) in FILES.items():
        path = os.path.join(ref_root, rel)
        blob = open(path, "rb").read()
        sha = hashlib.sha256(blob).hexdigest()
        if not sha.startswith(sha_prefix):
            raise SystemExit(f"{rel}: sha256 {sha[:16]} != expected {sha_prefix}; offsets untrusted")
        arrays = {}
        for key, (off, cnt, dt, shape, order) in fields.items():
            a = np.frombuffer(blob, dtype=dt, count=cnt, offset=off)
            arrays[key] = np.array(a.reshape(shape, order=order))
        arrays["_source"] = np.array(f"{rel} sha256={sha}")
        np.savez(os.path.join(out_dir, f"{name}.npz"), **arrays)
        print(f"{name}: {len(fields)} arrays from {rel}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    a = ap.parse_args()
    extract(a.ref, a.out)
