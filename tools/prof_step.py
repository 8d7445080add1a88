"""A few forward + adjoint steps of the headline workload for ncu:  python tools/prof_step.py [N] [steps]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import universal_differential_equations_b200 as ude
from helpers import glorot_theta, synthetic_ensemble
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
chain = ude.FastChain(ude.FastDense(2, 32, ude.tanh), ude.FastDense(32, 32, ude.tanh), ude.FastDense(32, 2))
s = ude.UDESolver(ude.LotkaVolterraUDE(chain), 0.0, 0.1, 30, 1, max_trajectories=N)
theta = glorot_theta((2, 32, 32, 2), seed=1)
u0, y = synthetic_ensemble(N)
s.set_params(torch.from_numpy(theta).cuda())
u0d, yd = torch.from_numpy(u0).cuda(), torch.from_numpy(y).cuda()
out = torch.empty((31, 2, N), device="cuda")
for _ in range(steps):
    s.forward(u0d, out=out); s.adjoint_l2(yd)
torch.cuda.synchronize()
