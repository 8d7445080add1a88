"""Config-4 shape (Nx = 256, 1-16-16-1): the same time span T = 0.2 with Tsit5 at its stability limit (dt = 1e-3 x 200 steps)
and with RKC2 (dt = 0.0125 x 16 steps, 8 stages): forward + adjoint times, and the difference of the two solutions / gradients."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta
Nx, N = 256, 8192
rng = np.random.default_rng(0)
layers = [ude.FastDense(1, 16, ude.tanh), ude.FastDense(16, 16, ude.tanh), ude.FastDense(16, 1)]
f = ude.FisherKPPUDE(ude.FastChain(*layers), Nx)
D0 = 0.01 * (Nx - 1) ** 2
theta = np.concatenate([glorot_theta((1, 16, 16, 1), seed=3), [1.0, -2.0, 1.0, 0.0, D0]]).astype(np.float32)
x = np.linspace(0, 1, Nx)
u0 = np.stack([0.5 * (np.tanh((x - (0.5 - d / 2)) / (d / 10)) - np.tanh((x - (0.5 + d / 2)) / (d / 10))) for d in rng.uniform(0.15, 0.5, N)], axis=1).astype(np.float32)
y = np.repeat(u0[None], 5, axis=0) * 0.9
th = torch.from_numpy(theta).cuda(); u0d = torch.from_numpy(u0).cuda(); yd = torch.from_numpy(y).cuda()
res = {}
for name, kw, dt, n in (("Tsit5", {}, 1e-3, 200), ("RKC2 s=8", {"alg": ude.ROCK2(eigen_est=4 * D0 + 2)}, 0.0125, 16)):
    s = ude.UDESolver(f, 0.0, dt, n, n // 4, max_trajectories=N, **kw)
    s.set_params(th)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(3):
        ev[0].record(); out = s.forward(u0d); ev[1].record(); L, g, _ = s.adjoint_l2(yd); ev[2].record(); torch.cuda.synchronize()
    res[name] = (out.cpu().numpy(), float(L), g.cpu().numpy())
    tf, ta = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print(f"{name}: fwd {tf:.2f} ms adj {ta:.2f} ms -> {N / ((tf + ta) * 1e-3):.3e} trajectories/s over T = 0.2")
    s.close()
a, b = res["Tsit5"], res["RKC2 s=8"]
print(f"RKC2 vs Tsit5: max |du| {np.abs(a[0] - b[0]).max():.2e}, loss rel {abs(a[1] - b[1]) / abs(a[1]):.2e}, grad rel {np.linalg.norm(a[2] - b[2]) / np.linalg.norm(a[2]):.2e}")
