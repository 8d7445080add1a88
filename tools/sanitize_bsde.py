"""Small invocations of the terminal-PDE kernels (fp64 fused sweeps, fp32 library path) and of the split host-buffer LV path, for
compute-sanitizer (memcheck / racecheck)."""
import math, sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta, synthetic_ensemble
d, hls = 100, 110
prob = ude.TerminalPDEProblem(ude.HJBTerminal(), ude.HJBNonlinearity(1.0), ude.ZeroDrift(), ude.ConstantDiffusion(math.sqrt(2.0)), np.zeros(d), (0.0, 1.0))
u0 = ude.Chain(ude.Dense(d, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, 1))
sg = ude.Chain(ude.Dense(d + 1, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, d))
alg = ude.NNPDENS(u0, sg, opt=ude.ADAM(0.03))
for dt in (torch.float64, torch.float32):
    s = ude.BSDESolver(prob, alg, 3, 77, dtype=dt)
    s.set_params(ude.initial_params_pde(alg))
    l, g, u = s.loss_gradient(77, 5)
    s.train_adam(ude.ADAM(0.03), 61, 3)
    print(dt, l, float(g.abs().sum()), u)
    s.close()
if len(sys.argv) > 1:
    N = 8200
    u0_, y = synthetic_ensemble(N)
    lv32 = ude.LotkaVolterraUDE(ude.FastChain(ude.FastDense(2, 32, ude.tanh), ude.FastDense(32, 32, ude.tanh), ude.FastDense(32, 2)))
    sv = ude.UDESolver(lv32, 0.0, 0.1, 6, 1, max_trajectories=N)
    gu = np.empty_like(u0_)
    print(sv.loss_gradient_host(glorot_theta((2, 32, 32, 2), seed=1), u0_, y[:7].copy(), grad_u0=gu)[0])
