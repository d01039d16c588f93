"""Trajectory sharding across ranks: no data-path collective, one all-gather of the result table at the end.

The reference has no distributed code; trajectories are independent given (complex, weights, seed),
so each rank runs a contiguous slice and NCCL (or gloo in the CPU tests) is used once per complex.
"""
import torch
import torch.distributed as dist


def rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(total, rank, world):
    """Contiguous, balanced slice [lo, hi) of `total` work items for `rank` (first total % world ranks get one more)."""
    if world <= 0 or rank < 0 or rank >= world:
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, rem = divmod(int(total), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, total, group=None):
    """All-gather row blocks laid out by shard_range -> [total, cols] on every rank (ragged-safe via padding)."""
    rank, world = rank_world(group)
    if world == 1:
        return local
    cols = local.shape[1]
    per = (total + world - 1) // world
    pad = torch.zeros(per, cols, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty(world * per, cols, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    rows = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        rows.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(rows, dim=0)
