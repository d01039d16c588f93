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
def _res_lig(size):
    """sizes[c] is either the number of residues N or the pair (R, L); without L a third of the complex is assumed."""
    if isinstance(size, (tuple, list)):
        return int(size[0]) + int(size[1]), int(size[1])
    return int(size), int(size) // 3


def complex_cost(size, num_samples):
    """Relative cost (in residue rows) of one launch sequence of `num_samples` trajectories of a complex (size = N or (R, L)):
    a fixed part per work item (the latency-bound kernels of every lock-step step, set_complex, the copy back) + the
    per-row work (edge / node kernels, 60 edges per residue; the ligand rows also pay the last layer and the coordinate
    head) with the O(N^2) stochastic-graph scan on top.  Fitted to the measured device times of all 25 db5 complexes at
    5 / 10 / 20 / 40 trajectories x 40 steps on one B200 (profiles/r02/c5_chunk_times.txt:
    time = 9.8 ms + 2.92 us per row-equivalent, mean error 3.4 %, worst 7.4 %)."""
    n, lig = _res_lig(size)
    return 3400.0 + float(num_samples) * (n + 0.3 * lig) * (1.0 + n / 6000.0)


def plan_work(sizes, num_samples, world, min_nodes=8192):
    """Assign every (complex, trajectory) of a set to a rank.

    Splitting each complex's trajectories evenly over all ranks balances perfectly but leaves every GPU with a handful
    of poses per launch (40 trajectories on 8 GPUs = 5 per GPU: the kernels run far below their throughput), while
    whole complexes per rank cannot balance a set whose largest member is 13x its smallest.  So: a complex is cut
    into at most `world` contiguous trajectory chunks, only as many as keep `min_nodes` residues per launch and only
    when its cost exceeds half a rank's share; chunks are placed longest-first on the least-loaded rank, each at the
    cost of its own launch sequence (complex_cost: the fixed part is paid per chunk), and the placement is then refined
    by moving or swapping single chunks off the most loaded rank for as long as that lowers the maximum load.
    sizes[c] = residues of complex c, or the pair (R, L).  Deterministic (every rank computes the same plan).  Returns a
    list of (complex index, lo, hi, rank) with the chunks of a complex in trajectory order.
    """
    if world <= 0:
        raise ValueError("world must be positive")
    costs = [complex_cost(n, num_samples) for n in sizes]
    share = sum(costs) / world if costs else 0.0
    chunks = []
    for c, size in enumerate(sizes):
        n = _res_lig(size)[0]
        by_size = max(1, (num_samples * n) // max(1, min_nodes))
        by_cost = int(-(-costs[c] // max(share / 2.0, 1e-9))) if share > 0 else 1      # ceil
        parts = max(1, min(world, num_samples, by_size, by_cost))
        for r in range(parts):
            lo, hi = shard_range(num_samples, r, parts)
            if hi > lo:
                chunks.append((c, lo, hi, complex_cost(size, hi - lo)))
    chunks.sort(key=lambda ch: (-ch[3], ch[0], ch[1]))
    load = [0.0] * world
    owner = []
    for c, lo, hi, w in chunks:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += w
        owner.append(r)
    # refinement: single moves / pairwise swaps off the most loaded rank (first improvement, fixed scan order)
    for _ in range(8 * len(chunks) + 8):
        hot = max(range(world), key=lambda k: (load[k], -k))
        done = True
        for i, (_, _, _, wi) in enumerate(chunks):
            if owner[i] != hot:
                continue
            for r in sorted(range(world), key=lambda k: (load[k], k)):
                if r == hot:
                    continue
                if max(load[r] + wi, load[hot] - wi) < load[hot] - 1e-9:                      # move i to r
                    load[hot] -= wi; load[r] += wi; owner[i] = r
                    done = False
                    break
                for j, (_, _, _, wj) in enumerate(chunks):
                    if owner[j] == r and wj < wi and max(load[r] + wi - wj, load[hot] - wi + wj) < load[hot] - 1e-9:
                        load[hot] += wj - wi; load[r] += wi - wj; owner[i], owner[j] = r, hot      # swap i and j
                        done = False
                        break
                if not done:
                    break
            if not done:
                break
        if done:
            break
    placed = sorted((c, lo, hi, owner[i]) for i, (c, lo, hi, _) in enumerate(chunks))
    return placed


def gather_objects(obj, group=None):
    """Python objects from every rank -> list on every rank (all_gather_object; results of a planned run are a few KB)."""
    rank, world = rank_world(group)
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out
