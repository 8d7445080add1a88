import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta, synthetic_ensemble
from oracle import oracle as O
chain = ude.FastChain(ude.FastDense(2, 32, ude.tanh), ude.FastDense(32, 32, ude.tanh), ude.FastDense(32, 2))
f = ude.LotkaVolterraUDE(chain)
theta = glorot_theta((2, 32, 32, 2), seed=1)
for N in (300, 65536):
    u0, y = synthetic_ensemble(N)
    solver = ude.UDESolver(f, 0.0, 0.1, 30, 1, max_trajectories=N)
    solver.set_params(torch.from_numpy(theta).cuda())
    u0d = torch.from_numpy(u0).cuda(); yd = torch.from_numpy(y).cuda()
    out = solver.forward(u0d); torch.cuda.synchronize()
    if N <= 4096:
        m = O.lv_model()
        l64, g64, gu64, out64 = O.ensemble_loss_grad(m, theta.astype(np.float64), u0, y, np.ones(2), 0.1, 30, want_out=True)
        o = out.cpu().numpy()
        print("N", N, "max|out-oracle|", np.abs(o - out64).max(), "rel viol", (np.abs(o - out64) / (1 + np.abs(out64))).max())
        loss, g, gu = solver.adjoint_l2(yd, want_grad_u0=True); torch.cuda.synchronize()
        print("  loss rel", abs(float(loss) - l64) / l64, "grad rel", np.linalg.norm(g.cpu().numpy() - g64) / np.linalg.norm(g64))
    else:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            solver.forward(u0d); solver.adjoint_l2(yd)
        tf = ta = 0
        for _ in range(10):
            e0.record(); solver.forward(u0d); e1.record(); solver.adjoint_l2(yd); e2.record(); torch.cuda.synchronize()
            tf += e0.elapsed_time(e1); ta += e1.elapsed_time(e2)
        print("N", N, "fwd ms", tf / 10, "adj ms", ta / 10, "traj/s", N / ((tf + ta) / 10 * 1e-3))
    solver.close()
