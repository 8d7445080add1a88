"""One forward + adjoint of the SEIR exposure UDE at one wave (for ncu captures)."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tools')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta
from seir_probe import make
N = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
u0, y = make(N)
chain = ude.FastChain(ude.FastDense(3, 64, ude.tanh), ude.FastDense(64, 64, ude.tanh), ude.FastDense(64, 1))
s = ude.UDESolver(ude.SEIRExposureUDE(chain), 0.0, 0.25, 84, 4, max_trajectories=N, loss_weights=[0, 1, 1, 1, 0, 0, 0])
s.set_params(torch.from_numpy(glorot_theta((3, 64, 64, 1), seed=2)).cuda())
s.forward(torch.from_numpy(u0).cuda()); s.adjoint_l2(torch.from_numpy(y).cuda()); torch.cuda.synchronize()
print("done")
