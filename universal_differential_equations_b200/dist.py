"""Multi-GPU plumbing: the ensemble shards across ranks, theta is replicated.

Trajectories are independent given theta; the only exchange on the path is the sum over
trajectories of dL/dtheta (and of L): ONE all-reduce of [grad_theta (P); loss (1)] per optimiser
step.  u0, data, saved states and dL/du0 never leave their GPU.
torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests) is the transport.
"""
import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of the ensemble owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_loss_grad(buf: torch.Tensor):
    """In-place sum over ranks of the packed [grad_theta; loss] buffer; no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf


class PeerAllReduce:
    """Sets up the fused reduce + all-reduce over NVLink peer memory for one UDESolver per rank
    (b200ude_peer_export / b200ude_peer_attach): CUDA IPC handles of the per-rank exchange buffers are gathered with
    torch.distributed, attached, and all ranks are synchronised once.  Afterwards `solver.adjoint_l2_allreduce(data)` returns
    the loss and grad_theta summed over all ranks with no NCCL call on the path.  world = 1 (no process group) works too."""

    def __init__(self, solver, group=None):
        self.solver = solver
        have_pg = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if have_pg else 0
        self.world = dist.get_world_size(group) if have_pg else 1
        mine = solver.peer_export()
        if self.world > 1:
            gathered = [None] * self.world
            dist.all_gather_object(gathered, mine, group=group)
        else:
            gathered = [mine]
        solver.peer_attach(self.rank, self.world, b"".join(gathered))
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=group)   # every rank's buffer is zeroed and mapped before anyone pushes

    def close(self):
        if self.world > 1:
            torch.cuda.synchronize()
            dist.barrier()
        self.solver.peer_detach()
