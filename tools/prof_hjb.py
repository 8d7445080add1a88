"""Per-kernel time table of one NNPDENS iteration (torch.profiler / CUPTI): python tools/prof_hjb.py [paths] [f32|f64]"""
import math, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import universal_differential_equations_b200 as ude
from torch.profiler import ProfilerActivity, profile
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dt = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else torch.float64
d, hls = 100, 110
prob = ude.TerminalPDEProblem(ude.HJBTerminal(), ude.HJBNonlinearity(1.0), ude.ZeroDrift(), ude.ConstantDiffusion(math.sqrt(2.0)), np.zeros(d), (0.0, 1.0))
u0 = ude.Chain(ude.Dense(d, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, 1))
sg = ude.Chain(ude.Dense(d + 1, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, d))
alg = ude.NNPDENS(u0, sg, opt=ude.ADAM(0.03))
s = ude.BSDESolver(prob, alg, 20, m, dtype=dt)
s.set_params(ude.initial_params_pde(alg))
s.train_adam(ude.ADAM(0.03), m, 5)
s.train_adam(ude.ADAM(0.03), m, 20); print("graph-replayed iteration: %.3f ms" % (s.last_train_ms() / 20))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    s.loss_gradient(m, 3); torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if str(getattr(e, "device_type", "")).endswith("CUDA"):
        a = agg.setdefault(e.name[:90], [0, 0.0]); a[0] += 1; a[1] += e.device_time
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us {v[0]:5d} x {100*v[1]/tot:5.1f}%  {k}")
print(f"sum of kernel times {tot/1e3:.3f} ms")
