"""One adaptive (abstol = reltol = 1e-6) forward + adjoint of the LV 2-32-32-2 chain on the tensor-core kernels (for ncu)."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta, synthetic_ensemble
N = 65536
chain = ude.FastChain(ude.FastDense(2, 32, ude.tanh), ude.FastDense(32, 32, ude.tanh), ude.FastDense(32, 2))
theta = glorot_theta((2, 32, 32, 2), seed=1)
u0, y = synthetic_ensemble(N)
s = ude.UDESolver(ude.LotkaVolterraUDE(chain), 0.0, 0.1, 30, 1, max_trajectories=N, adaptive=True, abstol=1e-6, reltol=1e-6, max_steps=128)
s.set_params(torch.from_numpy(theta).cuda())
st = torch.zeros(N, dtype=torch.int32, device="cuda")
s.forward(torch.from_numpy(u0).cuda(), status=st); s.adjoint_l2(torch.from_numpy(y).cuda()); torch.cuda.synchronize()
nacc = torch.empty(N, dtype=torch.int32)
print("status ok:", int((st == 0).sum()), "of", N)
