import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta, synthetic_ensemble
chain = ude.FastChain(ude.FastDense(2, 32, ude.tanh), ude.FastDense(32, 32, ude.tanh), ude.FastDense(32, 2))
f = ude.LotkaVolterraUDE(chain)
theta = glorot_theta((2, 32, 32, 2), seed=1)
for N in (1024, 65536):
    u0, y = synthetic_ensemble(N)
    s = ude.UDESolver(f, 0.0, 0.1, 30, 1, max_trajectories=N)
    th_h = torch.from_numpy(theta).pin_memory(); u0_h = torch.from_numpy(u0).pin_memory(); y_h = torch.from_numpy(y).pin_memory(); g_h = torch.empty(1218).pin_memory()
    for _ in range(3): s.loss_gradient_host(th_h, u0_h, y_h, grad_theta=g_h)
    t0 = time.perf_counter()
    for _ in range(20): s.loss_gradient_host(th_h, u0_h, y_h, grad_theta=g_h)
    dt = (time.perf_counter() - t0) / 20
    # pure H2D of data
    yd = torch.empty_like(y_h, device='cuda'); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): yd.copy_(y_h, non_blocking=True)
    torch.cuda.synchronize(); dc = (time.perf_counter() - t0) / 20
    print(f"N={N}: loss_gradient_host {dt*1e3:.3f} ms/call; H2D of data alone {dc*1e3:.3f} ms ({y.nbytes/dc/1e9:.1f} GB/s)")
    s.close()
