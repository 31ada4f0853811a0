"""Multi-GPU plumbing of the path: utterances are independent, so a batch is split into
contiguous per-rank blocks with NO data-path collective; the single exchange step is one all-gather of the
per-template u32 scores (and optionally best index / distance) at the end. torch.distributed is used only as
plumbing (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """contiguous block [lo, hi) of rank `rank` (blocks differ by at most one item)"""
    lo = n_items * rank // world
    hi = n_items * (rank + 1) // world
    return lo, hi


def gather_blocks(local, n_items, group=None):
    """all-gather row blocks of unequal length: `local` is this rank's [hi-lo, ...] tensor; returns the
    [n_items, ...] tensor on every rank. Uses ONE all_gather_into_tensor when the blocks are equal (the
    benchmark case), else pads to the largest block."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    rest = tuple(local.shape[1:])
    if len(set(sizes)) == 1:
        out = torch.empty((n_items,) + rest, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    m = max(sizes)
    pad = torch.zeros((m,) + rest, dtype=local.dtype, device=local.device)
    pad[: sizes[rank]] = local
    buf = torch.empty((world * m,) + rest, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * m: r * m + sizes[r]] for r in range(world)], 0)
