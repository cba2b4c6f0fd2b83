"""Data-parallel plumbing of the render path (one process per GPU, torch.distributed).

The path shards naturally: rendering shards rays / views, latent fitting shards
objects (no collective); training shards objects x rays and exchanges exactly one tensor per
step, the flat fp32 MLP gradient (714,756 floats = 2.86 MB), with a sum all-reduce.  Codes and
their gradients stay on the rank that owns the object.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, world_size, rank):
    """Contiguous, balanced [begin, end) share of n_items for `rank` (first ranks get the remainder)."""
    base, rem = divmod(int(n_items), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_mlp_grad(flat_grad, average=False):
    """Sum (or mean) the flat MLP gradient over all ranks in place.  No-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat_grad
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    if average:
        flat_grad.div_(dist.get_world_size())
    return flat_grad


def max_over_ranks(value, device):
    """Device-side timing convention: a multi-GPU number is the max over ranks."""
    t = torch.tensor([float(value)], device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
