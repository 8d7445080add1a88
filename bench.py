#!/usr/bin/env python3
"""bench.py -- ensemble trajectories/s, forward + interpolating adjoint, LV UDE (BASELINE config 2).

One "step" = one pass of the hot path over one batch of synthetic input: forward Tsit5 solve of
every trajectory (30 fixed steps of 0.1, states saved at every step) + InterpolatingAdjoint gradient
of the L2 trajectory-matching loss, summed over the ensemble (+ one NCCL all-reduce of
[grad_theta; loss] when N_gpus > 1).  Workload: 65 536 trajectories PER GPU (weak scaling; the
ensemble shards across ranks with no data-path collective), 2->32->32->2
tanh chain, Glorot theta (seed 1), u0 ~ U(0.2,1) x U(2,5) (seed 0), fp32.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
Under torchrun one rank per GPU.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_PER_GPU = 65536
N_STEPS, DT = 30, 0.1
WIDTHS = (2, 32, 32, 2)
P = 1218
METRIC = "ensemble trajectories/sec fwd+adjoint, LV UDE batch 65k"
UNIT = "trajectories/s"

# algorithmic cost per trajectory (DESIGN.md "Measurement")
FMA_RHS = 2 * 32 + 32 * 32 + 32 * 2                                  # 1152 FMA per chain evaluation
FLOP_FWD = 2.0 * FMA_RHS * (1 + 6 * N_STEPS)                           # 181 RHS evaluations
FLOP_ADJ = 2.0 * 3 * FMA_RHS * (6 * N_STEPS)                           # 180 backward stages x (fwd + J_u^T + J_theta^T)
BYTES_FWD = 4.0 * (2 + 2 * (N_STEPS + 1) * 2 + (6 * N_STEPS + 1) * 2)  # u0 + out + per-step store + dense output
BYTES_ADJ = 4.0 * ((N_STEPS + 1) * 2 * 2 + (6 * N_STEPS + 1) * 2 + 2)  # per-step store + data + dense output + grad_u0
FP32_PEAK_TFLOPS = 72.5  # measured FFMA/FFMA2 issue peak on this pool's B200 (profiles/r01_pipes_microbench.txt)


def synthetic(n, seed=0):
    from helpers import glorot_theta, synthetic_ensemble
    theta = glorot_theta(WIDTHS, seed=1)
    u0, y = synthetic_ensemble(n, n_steps=N_STEPS, dt=DT, seed=seed)
    return theta, u0, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.stop = [], threading.Event()
        self.cmd = ["nvidia-smi", f"--id={index}",
                    "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                    "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                    "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits"]
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                o = subprocess.run(self.cmd, capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline_run(theta, u0, y, sample, threads, reps=1):
    """The oracle port (C99/OpenMP, fp32, same algorithm) on a bounded sample of the same workload."""
    from oracle import oracle as O
    m = O.lv_model()
    th = theta.astype(np.float32)
    u0s, ys = np.ascontiguousarray(u0[:, :sample]), np.ascontiguousarray(y[:, :, :sample])
    w = np.ones(2, np.float32)
    best = float("inf")
    O.ensemble_loss_grad(m, th, u0s[:, :256], ys[:, :, :256], w, DT, N_STEPS, n_threads=threads, want_gu0=True)  # warm
    for _ in range(reps):
        t0 = time.perf_counter()
        O.ensemble_loss_grad(m, th, u0s, ys, w, DT, N_STEPS, n_threads=threads, want_gu0=True)
        best = min(best, time.perf_counter() - t0)
    return sample / best, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--cpu-sample", type=int, default=0, help="trajectories in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    if a.impl == "reference":
        # The reference's own path (OrdinaryDiffEq.jl + SciMLSensitivity.jl) needs Julia, which this image
        # does not have: the CPU arm is the oracle port, all host cores, bounded sample.
        if rank != 0:
            return
        theta, u0, y = synthetic(8192)
        sample = a.cpu_sample or 8192
        vals = []
        for _ in range(max(1, a.warmup)):
            cpu_baseline_run(theta, u0, y, min(sample, 1024), cores)
        t_all = 0.0
        for _ in range(a.steps):
            v, t = cpu_baseline_run(theta, u0, y, sample, cores)
            vals.append(v)
            t_all += t
        value = sample * a.steps / t_all
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * t_all / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "LV UDE ensemble, 2->32->32->2 tanh, Tsit5 dt=0.1 x30, fwd+InterpolatingAdjoint",
                       "sample_trajectories_per_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} trajectories per step x {a.steps} steps, oracle C99/OpenMP fp32, {cores} threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    import torch
    import torch.distributed as dist
    import universal_differential_equations_b200 as ude

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.n_per_gpu
    theta, u0, y = synthetic(n, seed=rank)   # every rank owns a different shard of the (world * n) ensemble
    chain = ude.FastChain(ude.FastDense(2, 32, ude.tanh), ude.FastDense(32, 32, ude.tanh), ude.FastDense(32, 2))
    f = ude.LotkaVolterraUDE(chain)
    solver = ude.UDESolver(f, 0.0, DT, N_STEPS, 1, max_trajectories=n, device=dev)
    th_d = torch.from_numpy(theta).to(dev)
    u0_d = torch.from_numpy(u0).to(dev)
    y_d = torch.from_numpy(y).to(dev)
    out_d = torch.empty((N_STEPS + 1, 2, n), device=dev)
    buf = torch.zeros(P + 1, device=dev)          # [grad_theta ; loss] -- the one all-reduced message
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    solver.set_params(th_d)
    # multi-GPU: the sum over ranks of [grad_theta; loss] runs inside the final reduction kernel over NVLink peer memory
    # (b200ude_adjoint_l2_allreduce); NCCL all-reduce only if the peer mapping cannot be set up on every rank
    peer = None
    if world > 1 and os.environ.get("B200UDE_PEER_ALLREDUCE", "1") != "0":
        ok = torch.ones(1, device=dev)
        try:
            peer = ude.PeerAllReduce(solver)
        except Exception as e:   # noqa: BLE001
            print(f"[bench] rank {rank}: peer-memory all-reduce unavailable ({e}); using NCCL", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            if peer is not None:
                solver.peer_detach()
            peer = None

    def step(ev=None, collective=True):
        solver.set_params(th_d)
        if ev:
            ev[0].record()
        solver.forward(u0_d, out=out_d)
        if ev:
            ev[1].record()
        if peer is not None and collective:
            solver.adjoint_l2_allreduce(y_d, want_grad_u0=False, grad_theta=buf[:P], loss=buf[P:])
            if ev:
                ev[2].record()
        else:
            solver.adjoint_l2(y_d, want_grad_u0=False, grad_theta=buf[:P], loss=buf[P:])
            if ev:
                ev[2].record()
            if collective:
                ude.allreduce_loss_grad(buf)
        if ev:
            ev[3].record()

    clk = ClockSampler(local)
    clk.__enter__()                                  # samples run from the warm-up through the timed region
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(a.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for i in range(a.steps):
        flush.zero_()                          # evict L2 between timed iterations (untimed)
        step(evs[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # the timed region of the default run lasts ~0.1 s, shorter than nvidia-smi's sampling period: keep the
    # same step running (untimed) until the sampler has seen the GPU under this load a few times
    t_probe = time.perf_counter()
    while rank == 0 and len(clk.rows) < 6 and time.perf_counter() - t_probe < 4.0:
        step(collective=False)                 # rank-local: no collective outside the lock-stepped region
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk.__exit__()
    clocks = clk.summary() if rank == 0 else None
    if clocks:
        clocks["window"] = "warm-up + timed steps + post-run probe of the same step (untimed)"
    t_step = [e[0].elapsed_time(e[3]) for e in evs]
    t_fwd = [e[0].elapsed_time(e[1]) for e in evs]
    t_adj = [e[1].elapsed_time(e[2]) for e in evs]
    total_ms = float(sum(t_step))
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t)
    value = world * n * a.steps / (total_ms * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call: pinned host inputs, H2D + kernels + D2H every step ----
    th_h = torch.from_numpy(theta).pin_memory()
    u0_h = torch.from_numpy(u0).pin_memory()
    y_h = torch.from_numpy(y).pin_memory()
    g_h = torch.empty(P).pin_memory()
    for _ in range(4):
        solver.loss_gradient_host(th_h, u0_h, y_h, grad_theta=g_h)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_steps = max(5, a.steps)
    for _ in range(e2e_steps):
        l_h, _, _ = solver.loss_gradient_host(th_h, u0_h, y_h, grad_theta=g_h)
        if world > 1:
            buf[:P].copy_(g_h, non_blocking=True)
            buf[P] = l_h
            ude.allreduce_loss_grad(buf)
            torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t)
    e2e_value = world * n * e2e_steps / e2e_s
    h2d = 4 * (P + u0.size + y.size)
    d2h = 4 * (P + 1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    adj_ms = float(np.mean(t_adj))
    fwd_ms = float(np.mean(t_fwd))
    roofline = {
        "kernel": "lv32::tc::adjoint_kernel (+ the ~5 us fixed-order reduce; events bracket both)",
        "bound": "hbm", "achieved": n * BYTES_ADJ / (adj_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
        "frac": n * BYTES_ADJ / (adj_ms * 1e-3) / 1e9 / hbm_peak,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
        "traffic": 132.5e6 * (n / 65536.0),
        "traffic_source": "ncu --set full capture of this kernel at N=65536: dram read 127.5 MB + write 5.0 MB (profiles/r01b_ncu_kernel_summaries.txt); algorithmic bytes %.1f MB" % (65536 * BYTES_ADJ / 1e6),
        "note": "the path is compute-bound by construction (SURVEY.md 8d, ~430 FLOP/B): roofline_fp32 (CUDA-core FP32 peak, which the tensor-core kernels bypass for the 32x32 layers) and roofline_xu (MUFU, the nearest bound of the tensor-core forward kernel) are the informative ones",
    }
    roofline_fp32 = {
        "kernel_adjoint": {"achieved": n * FLOP_ADJ / (adj_ms * 1e-3) / 1e12, "ms": adj_ms},
        "kernel_forward": {"achieved": n * FLOP_FWD / (fwd_ms * 1e-3) / 1e12, "ms": fwd_ms},
        "step": {"achieved": n * (FLOP_FWD + FLOP_ADJ) / (float(np.mean(t_step)) * 1e-3) / 1e12},
        "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
        "frac_adjoint": n * FLOP_ADJ / (adj_ms * 1e-3) / 1e12 / FP32_PEAK_TFLOPS,
        "frac_step": n * (FLOP_FWD + FLOP_ADJ) / (float(np.mean(t_step)) * 1e-3) / 1e12 / FP32_PEAK_TFLOPS,
        "peak_source": "measured FFMA2 issue peak, tools/microbench/pipes.cu on this pool (profiles/r01_pipes_microbench.txt)",
        "flop_per_trajectory": FLOP_FWD + FLOP_ADJ,
    }
    # MUFU (XU pipe) roofline: 2 MUFU per tanh, 64 tanh per chain evaluation; measured peak 16 lanes/clk/SM
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    xu_peak = 148 * 16 * sm_clock * 1e6
    roofline_xu = {
        "forward": {"achieved": n * 128.0 * (1 + 6 * N_STEPS) / (fwd_ms * 1e-3), "frac": n * 128.0 * (1 + 6 * N_STEPS) / (fwd_ms * 1e-3) / xu_peak},
        "adjoint": {"achieved": n * 128.0 * 6 * N_STEPS / (adj_ms * 1e-3), "frac": n * 128.0 * 6 * N_STEPS / (adj_ms * 1e-3) / xu_peak},
        "peak": xu_peak, "unit": "MUFU op/s",
        "peak_source": "16 MUFU lanes/clk/SM measured by tools/microbench/pipes.cu (profiles/r01_pipes_microbench.txt) x 148 SMs x sampled SM clock",
    }
    cpu = None
    if not a.no_cpu_baseline:
        sample = a.cpu_sample or 4096
        v, t = cpu_baseline_run(theta, u0, y, min(sample, n), cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{min(sample, n)} of the {n} trajectories, one pass ({t:.2f} s), oracle C99/OpenMP fp32 with {cores} threads"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LV UDE ensemble, 2->32->32->2 tanh, Tsit5 dt=0.1 x30 saveat 0.1, fwd + InterpolatingAdjoint + L2 loss",
                   "trajectories_per_gpu": n, "global_trajectories": world * n, "parallelism": f"ensemble-sharded x{world}",
                   "l2": "flushed between timed iterations (256 MiB memset, untimed)",
                   "timing": "CUDA events per step on the launch stream, summed over steps, max over ranks",
                   "allreduce": ("none (1 GPU)" if world == 1 else
                                 "fused into the final reduction kernel over NVLink peer memory (CUDA IPC, b200ude_adjoint_l2_allreduce)" if peer is not None
                                 else "NCCL all-reduce of [grad_theta; loss] (4.9 KB)")},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "note": "b200ude_loss_gradient_host: pinned host theta/u0/data -> H2D -> kernels -> D2H grad+loss, wall clock"},
        "gpu_launches": 3 * a.steps,
        "kernels_per_step": ["lv32::tc::forward_kernel", "lv32::tc::adjoint_kernel", "ude_reduce_exchange_kernel" if peer is not None else "ude_reduce_kernel"],
        "clocks": clocks, "roofline": roofline, "roofline_fp32": roofline_fp32, "roofline_xu": roofline_xu, "cpu_baseline": cpu,
        "kernel_ms": {"forward": fwd_ms, "adjoint_plus_reduce": adj_ms, "step": float(np.mean(t_step))},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
