"""One fp64 NNPDENS iteration at the benchmark's size (for ncu): python tools/hjb_once.py [paths]"""
import math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import universal_differential_equations_b200 as ude
m = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
d, hls = 100, 110
prob = ude.TerminalPDEProblem(ude.HJBTerminal(), ude.HJBNonlinearity(1.0), ude.ZeroDrift(), ude.ConstantDiffusion(math.sqrt(2.0)), np.zeros(d), (0.0, 1.0))
u0 = ude.Chain(ude.Dense(d, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, 1))
sg = ude.Chain(ude.Dense(d + 1, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, hls, ude.relu), ude.Dense(hls, d))
alg = ude.NNPDENS(u0, sg, opt=ude.ADAM(0.03))
s = ude.BSDESolver(prob, alg, 20, m, dtype=torch.float64)
s.set_params(ude.initial_params_pde(alg))
for i in range(3):
    s.loss_gradient(m, 3 + i)
torch.cuda.synchronize()
