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


# ---- many complexes (BASELINE config #5): work items are (complex, trajectory); plan them over the ranks -----------------
def complex_cost(n_res, num_samples):
    """Relative cost of `num_samples` trajectories of an n_res-residue complex: the per-node work (edge / node kernels,
    60 edges per residue) plus the O(N^2) stochastic-graph and energy scans, which take over above ~2000 residues
    (profiles/r01/launches_1N2C_v9.csv: 26 % of a step at N = 2548)."""
    return float(num_samples) * n_res * (1.0 + n_res / 6000.0)


def plan_work(sizes, num_samples, world, min_nodes=8192):
    """Assign every (complex, trajectory) of a set to a rank.

    Splitting each complex's trajectories evenly over all ranks balances perfectly but leaves every GPU with a handful
    of poses per launch (40 trajectories on 8 GPUs = 5 per GPU: the kernels run far below their throughput), while
    whole complexes per rank cannot balance a set whose largest member is 13x its smallest.  So: a complex is cut
    into at most `world` contiguous trajectory chunks, only as many as keep `min_nodes` residues per launch and only
    when its cost exceeds half a rank's share; chunks are then placed longest-first on the least-loaded rank.
    Deterministic (every rank computes the same plan).  Returns a list of (complex index, lo, hi, rank) with the chunks
    of a complex in trajectory order.
    """
    if world <= 0:
        raise ValueError("world must be positive")
    costs = [complex_cost(n, num_samples) for n in sizes]
    share = sum(costs) / world if costs else 0.0
    chunks = []
    for c, n in enumerate(sizes):
        by_size = max(1, (num_samples * n) // max(1, min_nodes))
        by_cost = int(-(-costs[c] // max(share / 2.0, 1e-9))) if share > 0 else 1      # ceil
        parts = max(1, min(world, num_samples, by_size, by_cost))
        for r in range(parts):
            lo, hi = shard_range(num_samples, r, parts)
            if hi > lo:
                chunks.append([c, lo, hi, costs[c] * (hi - lo) / num_samples])
    load = [0.0] * world
    placed = []
    for c, lo, hi, w in sorted(chunks, key=lambda ch: (-ch[3], ch[0], ch[1])):
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += w
        placed.append((c, lo, hi, r))
    placed.sort()
    return placed


def gather_objects(obj, group=None):
    """Python objects from every rank -> list on every rank (all_gather_object; results of a planned run are a few KB)."""
    rank, world = rank_world(group)
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out
