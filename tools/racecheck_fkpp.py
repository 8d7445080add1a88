"""Small Fisher-KPP (tuned packed / scalar) and 2-5-5-5-2 runs for compute-sanitizer --tool racecheck (shared-memory hazards of the
double-buffered neighbour exchange and of the staged gradient rows)."""
import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta
rng = np.random.default_rng(0)
for nx, n in ((26, 5), (27, 3)):
    fk = ude.FisherKPPUDE(ude.FastChain(ude.FastDense(1, 16, ude.tanh), ude.FastDense(16, 16, ude.tanh), ude.FastDense(16, 1)), nx)
    th = np.concatenate([glorot_theta((1, 16, 16, 1), seed=4), [1.1, -2.3, 0.9, 0.0, 0.01 * (nx - 1) ** 2]]).astype(np.float32)
    x = np.linspace(0, 1, nx)
    uf = np.stack([np.exp(-((x - 0.5) / d) ** 2) for d in rng.uniform(0.05, 0.3, n)], axis=1).astype(np.float32)
    yf = np.repeat(uf[None], 3, axis=0)
    s = ude.UDESolver(fk, 0.0, 2.5 / (16 * 0.01 * (nx - 1) ** 2), 4, 2, max_trajectories=n)
    s.set_params(torch.from_numpy(th).cuda()); s.forward(torch.from_numpy(uf).cuda()); L, g, _ = s.adjoint_l2(torch.from_numpy(yf).cuda())
    torch.cuda.synchronize(); print(nx, float(L)); s.close()
print("DONE")
