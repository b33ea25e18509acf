"""Range split of a batch over ranks and the three reductions a multi-GPU run needs.

PLUME items are independent, so rank g of G owns the contiguous range [g*n/G, (g+1)*n/G) and there is no
exchange step on the data path; torch.distributed (NCCL on GPUs, gloo in the CPU tests) is only used for
barriers, max-over-ranks timing and an AND over ranks of the correctness flags."""


def shard_range(n_total, rank, world):
    """Contiguous [first, last) of `n_total` items owned by `rank` out of `world` (sizes differ by at most 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    first = n_total * rank // world
    last = n_total * (rank + 1) // world
    return first, last


def reduce_max(values, dist=None, device=None):
    """Element-wise max over ranks of a list of floats (identity when not distributed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(values)
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def all_ranks_true(flag, dist=None, device=None):
    """AND over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return bool(flag)
    import torch
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t[0]))


def gather_counts(count, dist=None, device=None):
    """Per-rank item counts as a list (used to check that the shards tile the batch)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(count)]
    import torch
    t = torch.tensor([int(count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x[0]) for x in out]
