"""SEIR exposure UDE: tensor-core kernels (k_seir.cu) vs the runtime-shape kernels, parity + throughput."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta

def make(N, seed=9):
    rng = np.random.default_rng(seed)
    S0 = 14e6
    u0 = np.zeros((7, N), np.float32)
    u0[0] = 0.9 * S0 * rng.uniform(0.9, 1.0, N)
    u0[1:4] = rng.uniform(0, 50, (3, N)); u0[4] = S0; u0[5] = rng.uniform(0, 10, N); u0[6] = rng.uniform(0, 100, N)
    y = rng.uniform(0, 100, (22, 7, N)).astype(np.float32)
    return u0, y

def run(tc, N, theta, u0, y, reps=3):
    os.environ["B200UDE_SEIR_TC"] = "1" if tc else "0"
    chain = ude.FastChain(ude.FastDense(3, 64, ude.tanh), ude.FastDense(64, 64, ude.tanh), ude.FastDense(64, 1))
    s = ude.UDESolver(ude.SEIRExposureUDE(chain), 0.0, 0.25, 84, 4, max_trajectories=N, loss_weights=[0, 1, 1, 1, 0, 0, 0])
    th = torch.from_numpy(theta).cuda(); u0d = torch.from_numpy(u0).cuda(); yd = torch.from_numpy(y).cuda()
    s.set_params(th)
    out = s.forward(u0d); loss, g, gu = s.adjoint_l2(yd, want_grad_u0=True); torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = ta = 0.0
    for _ in range(reps):
        e[0].record(); s.forward(u0d); e[1].record(); s.adjoint_l2(yd); e[2].record(); torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]) / reps; ta += e[1].elapsed_time(e[2]) / reps
    r = (out.cpu().numpy(), float(loss), g.cpu().numpy(), gu.cpu().numpy(), tf, ta)
    s.close()
    return r

if __name__ == '__main__':
  theta = glorot_theta((3, 64, 64, 1), seed=2)
  for N in (300, int(sys.argv[1]) if len(sys.argv) > 1 else 37888):
      u0, y = make(N)
      a = run(True, N, theta, u0, y)
      b = run(False, N, theta, u0, y, reps=1)
      sc = np.abs(b[0]).max(axis=(0, 2), keepdims=True)
      print(f"N={N}: out err/scale {np.max(np.abs(a[0] - b[0]) / (sc + 1e-3)):.2e}  loss {a[1]:.6e} vs {b[1]:.6e}  "
            f"grad rel {np.linalg.norm(a[2] - b[2]) / np.linalg.norm(b[2]):.2e}  gu0 rel {np.linalg.norm(a[3] - b[3]) / np.linalg.norm(b[3]):.2e}")
      print(f"   TC: fwd {a[4]:.3f} ms adj {a[5]:.3f} ms -> {N / ((a[4] + a[5]) * 1e-3):.3e} traj/s   generic: fwd {b[4]:.2f} adj {b[5]:.2f} ms -> {N / ((b[4] + b[5]) * 1e-3):.3e} traj/s")
