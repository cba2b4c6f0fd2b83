"""Data-parallel plumbing of the render path (one process per GPU, torch.distributed over NCCL / NVLink).

The path shards naturally:

* rendering shards (object, view) pairs and latent fitting shards test objects -- `shard_range`, no collective;
* training shards the objects of a step; the ranks exchange exactly one tensor per step, the flat fp32 MLP
  gradient (714,756 floats = 2.86 MB), with a sum all-reduce (`allreduce_mlp_grad`, asynchronous so that the
  per-object code updates overlap it).  Code rows and their gradients stay on the rank that owns the object;
  `gather_owned_rows` assembles the full tables on every rank for a checkpoint.

Used by `Trainer`, `CodeFitter.fit_batch`, `render_dataset` and bench.py; `tests/test_dist_gloo.py` covers the
world_size-2 logic on CPU, `tests/test_gpu_dist.py` on two GPUs.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group; (0, 1) without one."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, world_size, rank):
    """Contiguous, balanced [begin, end) share of n_items for `rank` (first ranks get the remainder)."""
    base, rem = divmod(int(n_items), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class _Done:
    def wait(self):
        return True


def allreduce_mlp_grad(flat_grad, average=False, async_op=False):
    """Sum (or mean) the flat MLP gradient over all ranks in place.  With `async_op` the collective runs on NCCL's
    own stream and a handle is returned: call `.wait()` before the optimiser reads the gradient.  No-op (and a
    completed handle) without a process group."""
    if world()[1] == 1:
        return _Done() if async_op else flat_grad
    if average:
        flat_grad.div_(world()[1])
    work = dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, async_op=async_op)
    return work if async_op else flat_grad


def gather_owned_rows(table, owner_of_row):
    """Every rank holds a full [n, d] table but only the rows it owns are current (`owner_of_row[i]` = rank).
    Returns the table with every row taken from its owner (a masked sum all-reduce: checkpoints only)."""
    rank, ws = world()
    if ws == 1:
        return table.clone()
    mine = (torch.as_tensor(owner_of_row, device=table.device) == rank).to(table.dtype).unsqueeze(1)
    out = table.detach() * mine
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def gather_varlen(local, counts=None):
    """Concatenate per-rank [n_r, ...] tensors (n_r may differ) in rank order on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    n = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    ns = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(ns, n)
    ns = [int(t.item()) for t in ns]
    m = max(ns)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:k] for b, k in zip(bufs, ns)], 0)


def max_over_ranks(value, device):
    """Device-side timing convention: a multi-GPU number is the max over ranks."""
    t = torch.tensor([float(value)], device=device)
    if world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
