"""Gradient accuracy of the LV32 adjoint variants against the fp64 oracle (and of the SEIR variants against each other)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta, synthetic_ensemble
from oracle import oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
theta = glorot_theta((2, 32, 32, 2), seed=1)
u0, y = synthetic_ensemble(N)
m = O.lv_model()
l64, g64, _ = O.ensemble_loss_grad(m, theta.astype(np.float64), u0, y, np.ones(2), 0.1, 30)
chain = ude.FastChain(ude.FastDense(2, 32, ude.tanh), ude.FastDense(32, 32, ude.tanh), ude.FastDense(32, 2))
W2 = slice(96, 96 + 1024)
for v in ("1", "2", "0"):
    os.environ["B200UDE_ADJ_TC"] = v
    s = ude.UDESolver(ude.LotkaVolterraUDE(chain), 0.0, 0.1, 30, 1, max_trajectories=N)
    s.set_params(torch.from_numpy(theta).cuda()); s.forward(torch.from_numpy(u0).cuda())
    L, g, _ = s.adjoint_l2(torch.from_numpy(y).cuda()); g = g.cpu().numpy().astype(np.float64)
    print(f"LV32 N={N} ADJ_TC={v}: |g-g64|/|g64| all {np.linalg.norm(g - g64) / np.linalg.norm(g64):.3e}  W2 block {np.linalg.norm(g[W2] - g64[W2]) / np.linalg.norm(g64[W2]):.3e}  max-abs-rel {np.abs(g - g64).max() / np.abs(g64).max():.3e}")
    s.close()
