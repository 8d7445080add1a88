"""torchrun --nproc-per-node N tools/peer_allreduce_check.py: the fused reduce + all-reduce over NVLink peer memory against
NCCL all-reduce of the per-rank results (values and timing)."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import universal_differential_equations_b200 as ude
from helpers import glorot_theta, synthetic_ensemble
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 65536
theta = glorot_theta((2, 32, 32, 2), seed=1)
u0, y = synthetic_ensemble(N, seed=rank)
chain = ude.FastChain(ude.FastDense(2, 32, ude.tanh), ude.FastDense(32, 32, ude.tanh), ude.FastDense(32, 2))
s = ude.UDESolver(ude.LotkaVolterraUDE(chain), 0.0, 0.1, 30, 1, max_trajectories=N)
s.set_params(torch.from_numpy(theta).cuda())
u0d, yd = torch.from_numpy(u0).cuda(), torch.from_numpy(y).cuda()
pa = ude.PeerAllReduce(s)
buf = torch.empty(s.P + 1, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
t_nccl = t_peer = 0.0
ok = True
for it in range(30):
    s.forward(u0d)
    ev[0].record()
    L, g, _ = s.adjoint_l2(yd, grad_theta=buf[:s.P], loss=buf[s.P:])
    dist.all_reduce(buf)
    ev[1].record()
    Lp, gp, _ = s.adjoint_l2_allreduce(yd)
    ev[2].record()
    torch.cuda.synchronize()
    ref = buf.clone()
    rel = float((gp - ref[:s.P]).norm() / ref[:s.P].norm())
    ok = ok and rel < 1e-6 and abs(float(Lp) - float(ref[s.P])) <= 1e-6 * abs(float(ref[s.P]))
    # every rank must hold bitwise identical sums
    chk = torch.cat([gp, Lp]).clone()
    lst = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    ok = ok and all(torch.equal(lst[0], x) for x in lst)
    if it >= 10:
        t_nccl += ev[0].elapsed_time(ev[1]) / 20; t_peer += ev[1].elapsed_time(ev[2]) / 20
if rank == 0:
    print(f"world={world}: peer-memory fused all-reduce {'OK' if ok else 'MISMATCH'} (last rel diff vs NCCL {rel:.2e}); "
          f"adjoint+reduce+NCCL {t_nccl:.4f} ms, adjoint+fused reduce/all-reduce {t_peer:.4f} ms")
# multi-GPU training loop entirely on the devices: every rank must end with bitwise identical parameters, equal to a
# host-driven loop that all-reduces with NCCL
s.set_params(torch.from_numpy(theta).cuda()); s.adam_reset()
hist = s.train_adam(ude.ADAM(0.01), u0d, yd, 8)
th_dev = s.get_params().clone()
lst = [torch.empty_like(th_dev) for _ in range(world)]
dist.all_gather(lst, th_dev)
same = all(torch.equal(lst[0], x) for x in lst)
th = torch.from_numpy(theta).cuda(); m = torch.zeros_like(th); v = torch.zeros_like(th); ref_hist = []
for it in range(1, 9):
    s.set_params(th); s.forward(u0d)
    L, g, _ = s.adjoint_l2(yd, grad_theta=buf[:s.P], loss=buf[s.P:])
    dist.all_reduce(buf)
    ref_hist.append(float(buf[s.P])); g = buf[:s.P].clone()
    m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
    th = th - 0.01 * (m / (1 - 0.9 ** it)) / (torch.sqrt(v / (1 - 0.999 ** it)) + 1e-8)
rel = float((th_dev - th).norm() / th.norm())
if rank == 0:
    print(f"world={world}: on-device multi-GPU ADAM loop: replicas identical={same}, theta rel diff vs host-driven NCCL loop {rel:.2e}, "
          f"loss history rel diff {max(abs(a - b) / abs(b) for a, b in zip(hist.cpu().tolist(), ref_hist)):.2e}")
pa.close(); s.close()
dist.destroy_process_group()
