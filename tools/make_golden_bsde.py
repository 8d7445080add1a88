"""Writes tests/golden/hjb_small.npz: a frozen input/output vector of oracle/bsde_oracle.py (one NNPDENS iteration of the
highdim_pde/lambaem.jl problem at d = 12, hls = 16, 24 paths, 5 Euler-Maruyama steps; a second case at the script's tie point x0 = 0 with
zero biases).  There is no reference artefact for this path (its only check is `@test error_l2 < 0.2`), so this fixture pins the
restatement against regressions and gives the GPU tests a target that does not depend on the oracle's code at run time.
    python tools/make_golden_bsde.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bsde_oracle as bo

d, hls, M, N = 12, 16, 24, 5
rng = np.random.default_rng(2026)
x0 = 0.3 * rng.standard_normal(d)
theta = bo.init_params(d, hls, seed=7) + 0.05 * rng.standard_normal(sum(bo.num_params(d, hls)))
loss, grad, u0 = bo.loss_and_grad(theta, d, hls, x0, 1.0, N, M, seed=41)
theta_t = bo.init_params(d, hls, seed=8)
loss_t, grad_t, u0_t = bo.loss_and_grad(theta_t, d, hls, np.zeros(d), 1.0, N, M, seed=42)
np.savez(os.path.join(ROOT, "tests", "golden", "hjb_small.npz"), d=d, hls=hls, M=M, N=N, x0=x0, theta=theta, loss=loss, grad=grad, u0=u0,
         seed=41, theta_tie=theta_t, loss_tie=loss_t, grad_tie=grad_t, u0_tie=u0_t, seed_tie=42,
         normals_step3=bo.normals(41, 3, 8, d))
print("wrote hjb_small.npz: loss", loss, "tie loss", loss_t)
