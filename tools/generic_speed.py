"""Throughput of the runtime-shape (generic) kernel family on representative shapes (reported in DESIGN.md)."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta, synthetic_ensemble

def timeit(solver, theta, u0, y, reps=3):
    th = torch.from_numpy(theta).cuda(); u0d = torch.from_numpy(u0).cuda(); yd = torch.from_numpy(y).cuda()
    solver.set_params(th)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    solver.forward(u0d); solver.adjoint_l2(yd); torch.cuda.synchronize()
    tf = ta = 0.0
    for _ in range(reps):
        e[0].record(); out = solver.forward(u0d); e[1].record(); solver.adjoint_l2(yd); e[2].record(); torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); ta += e[1].elapsed_time(e[2])
    return tf / reps, ta / reps

rng = np.random.default_rng(0)
# SEIR exposure UDE, 3-64-64-1, Tsit5 dt=0.25 x84, saved daily (seir_exposure.jl shapes)
N = 8192
chain = ude.FastChain(ude.FastDense(3, 64, ude.tanh), ude.FastDense(64, 64, ude.tanh), ude.FastDense(64, 1))
f = ude.SEIRExposureUDE(chain)
theta = glorot_theta((3, 64, 64, 1), seed=2)
S0 = 14e6
u0 = np.zeros((7, N), np.float32); u0[0] = 0.9 * S0; u0[1:4] = rng.uniform(0, 50, (3, N)); u0[4] = S0
y = rng.uniform(0, 100, (22, 7, N)).astype(np.float32)
s = ude.UDESolver(f, 0.0, 0.25, 84, 4, max_trajectories=N, loss_weights=[0, 1, 1, 1, 0, 0, 0])
tf, ta = timeit(s, theta, u0, y)
print(f"SEIR 3-64-64-1 generic, N={N}: fwd {tf:.2f} ms adj {ta:.2f} ms -> {N / ((tf + ta) * 1e-3):.3e} traj/s")
s.close()
# Fisher-KPP 26-point grid, 1-10-20-10-1
N = 4096
layers = [ude.FastDense(1, 10, ude.tanh), ude.FastDense(10, 20, ude.tanh), ude.FastDense(20, 10, ude.tanh), ude.FastDense(10, 1)]
f = ude.FisherKPPUDE(ude.FastChain(*layers), 26)
theta = np.concatenate([glorot_theta((1, 10, 20, 10, 1), seed=3), [1.1, -2.5, 1.0, 0.0, 6.5]]).astype(np.float32)
x = np.linspace(0, 1, 26)
u0 = np.stack([0.5 * (np.tanh((x - (0.5 - d / 2)) / (d / 10)) - np.tanh((x - (0.5 + d / 2)) / (d / 10))) for d in rng.uniform(0.15, 0.5, N)], axis=1).astype(np.float32)
y = np.repeat(u0[None], 11, axis=0)
s = ude.UDESolver(f, 0.0, 0.0125, 400, 40, max_trajectories=N)
tf, ta = timeit(s, theta, u0, y, reps=2)
print(f"FKPP Nx=26 1-10-20-10-1 generic, N={N}, 400 steps: fwd {tf:.2f} ms adj {ta:.2f} ms -> {N / ((tf + ta) * 1e-3):.3e} traj/s")
s.close()
# LV 2-32-32-2 adaptive (generic kernels), tol 1e-6
N = 65536
chain = ude.FastChain(ude.FastDense(2, 32, ude.tanh), ude.FastDense(32, 32, ude.tanh), ude.FastDense(32, 2))
f = ude.LotkaVolterraUDE(chain)
theta = glorot_theta((2, 32, 32, 2), seed=1)
u0, y = synthetic_ensemble(N)
s = ude.UDESolver(f, 0.0, 0.1, 30, 1, max_trajectories=N, adaptive=True, abstol=1e-6, reltol=1e-6, max_steps=128)
tf, ta = timeit(s, theta, u0, y, reps=2)
print(f"LV 2-32-32-2 ADAPTIVE tol 1e-6 generic, N={N}: fwd {tf:.2f} ms adj {ta:.2f} ms -> {N / ((tf + ta) * 1e-3):.3e} traj/s")
s.close()
