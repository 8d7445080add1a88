"""One forward + adjoint of the tuned Fisher-KPP kernels (for ncu captures)."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta
Nx, N = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 1184
rng = np.random.default_rng(0)
layers = [ude.FastDense(1, 16, ude.tanh), ude.FastDense(16, 16, ude.tanh), ude.FastDense(16, 1)]
f = ude.FisherKPPUDE(ude.FastChain(*layers), Nx)
theta = np.concatenate([glorot_theta((1, 16, 16, 1), seed=3), [1.0, -2.0, 1.0, 0.0, 0.01 * (Nx - 1) ** 2]]).astype(np.float32)
x = np.linspace(0, 1, Nx)
u0 = np.stack([0.5 * (np.tanh((x - (0.5 - d / 2)) / (d / 10)) - np.tanh((x - (0.5 + d / 2)) / (d / 10))) for d in rng.uniform(0.15, 0.5, N)], axis=1).astype(np.float32)
y = np.repeat(u0[None], 11, axis=0)
s = ude.UDESolver(f, 0.0, 1e-3, 100, 10, max_trajectories=N)
s.set_params(torch.from_numpy(theta).cuda())
s.forward(torch.from_numpy(u0).cuda()); s.adjoint_l2(torch.from_numpy(y).cuda()); torch.cuda.synchronize()
print("done")
