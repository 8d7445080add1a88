"""Writes tests/golden/hjb_script_curve.npz: the oracle's training curve at the reference script's exact configuration
(highdim_pde/lambaem.jl: d = 100, hls = 110, m = 100 trajectories, ADAM(0.03), 20 Euler-Maruyama steps), 300 iterations from
init_params(seed 0), Brownian seeds 1, 2, ...  -- loss and u0(x0) before every update.   python tools/make_golden_bsde_curve.py (~1 min)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bsde_oracle as bo
d, hls, M, N, iters = 100, 110, 100, 20, 300
theta0 = bo.init_params(d, hls, seed=0)
t0 = time.time()
theta, losses, u0s = bo.train(theta0, d, hls, np.zeros(d), 1.0, N, M, iters, eta=0.03, seed0=1)
np.savez(os.path.join(ROOT, "tests", "golden", "hjb_script_curve.npz"), d=d, hls=hls, M=M, N=N, iters=iters, eta=0.03, seed0=1, init_seed=0,
         losses=losses, u0s=u0s, theta_final_norm=np.linalg.norm(theta))
print(f"{time.time() - t0:.1f} s; loss {losses[0]:.4f} -> {losses[-1]:.4f}; u0 {u0s[0]:.4f} -> {u0s[-1]:.4f}")
