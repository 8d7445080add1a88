"""Optimiser-iteration rate: on-device ADAM loop (b200ude_train_adam, CUDA graph) vs the host-driven loop, on the
reference's own training shape (scenario_1.jl: one trajectory, 2-5-5-5-2 rbf chain, 31 save points) and on the
BASELINE ensemble (65 536 trajectories, 2-32-32-2)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta, synthetic_ensemble

def rate(solver, theta, u0d, yd, iters, graph):
    os.environ["B200UDE_TRAIN_GRAPH"] = "1" if graph else "0"
    solver.set_params(torch.from_numpy(theta).cuda()); solver.adam_reset()
    solver.train_adam(ude.ADAM(0.01), u0d, yd, 5); torch.cuda.synchronize()
    t = time.perf_counter(); l = solver.train_adam(ude.ADAM(0.01), u0d, yd, iters); torch.cuda.synchronize()
    return iters / (time.perf_counter() - t), float(l[-1])

def host_rate(solver, theta, u0d, yd, iters):
    th = torch.from_numpy(theta).cuda(); m = torch.zeros_like(th); v = torch.zeros_like(th)
    torch.cuda.synchronize(); t = time.perf_counter()
    for it in range(1, iters + 1):
        solver.set_params(th); solver.forward(u0d); L, g, _ = solver.adjoint_l2(yd)
        l = float(L)   # the reference's callback reads the loss every iteration
        m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
        th = th - 0.01 * (m / (1 - 0.9 ** it)) / (torch.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
    torch.cuda.synchronize()
    return iters / (time.perf_counter() - t)

chain5 = ude.FastChain(ude.FastDense(2, 5, ude.rbf), ude.FastDense(5, 5, ude.rbf), ude.FastDense(5, 5, ude.rbf), ude.FastDense(5, 2))
s = ude.UDESolver(ude.LotkaVolterraUDE(chain5), 0.0, 0.1 / 4, 120, 4, max_trajectories=1)
theta = glorot_theta((2, 5, 5, 5, 2), seed=0)
u0, y = synthetic_ensemble(1)
u0d, yd = torch.from_numpy(u0).cuda(), torch.from_numpy(y).cuda()
print("scenario_1 shape (N=1, 120 steps):  graph %.0f it/s  eager %.0f it/s  host-driven %.0f it/s" %
      (rate(s, theta, u0d, yd, 2000, True)[0], rate(s, theta, u0d, yd, 2000, False)[0], host_rate(s, theta, u0d, yd, 500)))
s.close()
N = 65536
chain = ude.FastChain(ude.FastDense(2, 32, ude.tanh), ude.FastDense(32, 32, ude.tanh), ude.FastDense(32, 2))
s = ude.UDESolver(ude.LotkaVolterraUDE(chain), 0.0, 0.1, 30, 1, max_trajectories=N)
theta = glorot_theta((2, 32, 32, 2), seed=1)
u0, y = synthetic_ensemble(N)
u0d, yd = torch.from_numpy(u0).cuda(), torch.from_numpy(y).cuda()
rg, re_, rh = rate(s, theta, u0d, yd, 200, True)[0], rate(s, theta, u0d, yd, 200, False)[0], host_rate(s, theta, u0d, yd, 100)
print("BASELINE ensemble (N=65536): graph %.1f it/s (%.2f M traj/s)  eager %.1f it/s  host-driven %.1f it/s" % (rg, rg * N / 1e6, re_, rh))
s.close()
