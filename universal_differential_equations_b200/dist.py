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
