import os, sys, numpy as np, torch, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import universal_differential_equations_b200 as ude
ann = ude.FastChain(ude.FastDense(3, 64, ude.tanh), ude.FastDense(64, 64, ude.tanh), ude.FastDense(64, 1))
p = ude.initial_params(ann, np.random.default_rng(0))
u0 = np.array([14e6 - 1e3, 0.0, 100.0, 0.0, 14e6, 0.0, 0.0], np.float32)[:, None]
for alg in (ude.Vern7(), ude.Tsit5()):
    for tol in (1e-4, 1e-3, 1e-5):
        for ms in (512, 4096):
            s = ude.UDESolver(ude.SEIRExposureUDE(ann), 0.0, 1.0, 21, 1, max_trajectories=1, alg=alg, adaptive=True, abstol=tol, reltol=tol, max_steps=ms,
                              loss_weights=[0, 1, 1, 1, 0, 0, 0])
            s.set_params(torch.from_numpy(p).cuda())
            st = torch.zeros(1, dtype=torch.int32, device="cuda")
            out = s.forward(torch.from_numpy(u0).cuda(), status=st)
            torch.cuda.synchronize()
            print(type(alg).__name__, tol, ms, "status", int(st[0]), "last", out[-1, :, 0].cpu().numpy())
            s.close()
