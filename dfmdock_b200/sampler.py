"""Reverse-diffusion samplers.

Euler_Maruyama_sampler keeps the reference signature and RNG consumption order
(src/inference_base.py:390-468; centre_mode=1 gives the src/inference.py:213-370 variant):
numpy normal(4) [scipy Rotation.random] -> torch normal(1,3) -> per step {Exp(1) [N,N-20], randn(1,3), randn(1,3)}
-> final forward {Exp(1)}.  Every tensor op of the loop body runs in the CUDA library.

sample_trajectories is the batched form the reference lacks: all trajectories of a complex advance in lock step
on one GPU (Philox noise), optionally sharded over ranks with one gather at the end (dfmdock_b200.distributed).
"""
import torch

from . import distributed as dist_utils


def Euler_Maruyama_sampler(model, batch, num_steps=40, device="cpu", batch_size=1, eps=1e-3, use_clash_force=False,
                           noise_annealing=False, tr_noise_scale=0.5, rot_noise_scale=0.5, centre_mode=0, ode=False,
                           trajectory=None):
    """ode / trajectory are the two extras of the class-form sampler (src/inference_mlsb.py:264-350): the
    probability-flow branch of torch_reverse, and a list that receives lig_pos after the random init and every step."""
    from scipy.spatial.transform import Rotation

    dev = model.device
    time_steps = torch.linspace(1.0, eps, num_steps, device=dev)
    dt = time_steps[0] - time_steps[1]
    ts_host = time_steps.tolist()
    dt_host = float(dt)

    rec_pos = batch["rec_pos"].clone().to(dev)
    lig_pos0 = batch["lig_pos"].clone().to(dev)
    model.set_complex(batch)

    # randomize_pose (inference_base.py:318-340): same draws, same order
    rot0 = torch.from_numpy(Rotation.random().as_matrix()).float()
    tr0 = torch.normal(0.0, 30.0, size=(1, 3), device=dev)
    lig_pos, tr_update, rot_update = model.randomize_pose(lig_pos0, 1, rot0=rot0[None], tr0=tr0, centre_mode=centre_mode)

    if trajectory is not None:
        trajectory.append(lig_pos[0].clone())
    output = None
    with torch.no_grad():
        for i in range(num_steps):
            t_host = ts_host[i]
            is_last = i == num_steps - 1
            batch["t"] = torch.ones(batch_size, device=dev) * t_host
            batch["rec_pos"] = rec_pos
            batch["lig_pos"] = lig_pos[0]
            output = model(batch)
            if noise_annealing:
                ns_tr = ns_rot = t_host
            elif is_last:
                ns_tr = ns_rot = 0.0
            else:
                ns_tr, ns_rot = tr_noise_scale, rot_noise_scale
            # torch_reverse draws randn(1,3) only on the SDE branch (so3_diffuser.py:362-367): the ODE branch consumes no RNG
            z = None if ode else torch.cat([torch.randn(1, 3, device=dev), torch.randn(1, 3, device=dev)], dim=0)[None]   # rot, then tr
            model.reverse_step(lig_pos, rot_update, tr_update, output["tr_score"], output["rot_score"], t_host, dt_host,
                               ns_rot, ns_tr, z=z, use_clash_force=use_clash_force, centre_mode=centre_mode, ode=ode)
            if trajectory is not None:
                trajectory.append(lig_pos[0].clone())
            if is_last:
                batch["rec_pos"] = rec_pos
                batch["lig_pos"] = lig_pos[0]
                output = model(batch)
    return rec_pos, lig_pos[0], rot_update, tr_update, output


def sample_trajectories(model, batch, num_samples, num_steps=40, eps=1e-3, use_clash_force=False, noise_annealing=False,
                        tr_noise_scale=0.5, rot_noise_scale=0.5, centre_mode=0, seed=0, max_batch=None, group=None,
                        gather_poses=False, ode=False):
    """All `num_samples` trajectories of one complex.  Under torch.distributed each rank runs a contiguous slice
    (trajectory k always uses Philox subsequence k, so the result does not depend on the number of ranks) and the
    [T,8] result table (rot_update 3, tr_update 3, energy, num_clashes) is all-gathered once.

    Returns dict with global tensors: rot_update [T,3], tr_update [T,3], energy [T], num_clashes [T], best (argmin energy),
    and lig_pos [T,L,3,3] if gather_poses (else the local slice under "lig_pos_local").
    """
    rank, world = dist_utils.rank_world(group)
    lo, hi = dist_utils.shard_range(num_samples, rank, world)
    model.set_complex(batch)
    dev = model.device
    L = batch["lig_pos"].shape[0]
    chunks = []
    step = max_batch or max(hi - lo, 1)
    for c0 in range(lo, hi, step):
        c1 = min(hi, c0 + step)
        chunks.append(model.sample(batch["lig_pos"], c1 - c0, num_steps=num_steps, eps=eps, tr_noise_scale=tr_noise_scale,
                                   rot_noise_scale=rot_noise_scale, use_clash_force=use_clash_force,
                                   noise_annealing=noise_annealing, centre_mode=centre_mode, seed=seed, stream_base=c0,
                                   ode=ode))
    if chunks:
        local = {k: torch.cat([c[k] for c in chunks], dim=0) for k in chunks[0]}
    else:
        local = {"lig_pos": torch.empty(0, L, 3, 3, device=dev), "rot_update": torch.empty(0, 3, device=dev),
                 "tr_update": torch.empty(0, 3, device=dev), "energy": torch.empty(0, device=dev),
                 "num_clashes": torch.empty(0, dtype=torch.int32, device=dev)}
    table = torch.cat([local["rot_update"], local["tr_update"], local["energy"][:, None],
                       local["num_clashes"].float()[:, None]], dim=1)
    full = dist_utils.gather_rows(table, num_samples, group)
    out = {"rot_update": full[:, 0:3], "tr_update": full[:, 3:6], "energy": full[:, 6],
           "num_clashes": full[:, 7].round().long(), "lig_pos_local": local["lig_pos"], "local_range": (lo, hi)}
    if gather_poses:
        out["lig_pos"] = dist_utils.gather_rows(local["lig_pos"].reshape(hi - lo, L * 9), num_samples, group).view(-1, L, 3, 3)
    out["best"] = int(torch.argmin(out["energy"]).item()) if num_samples > 0 else -1
    return out


_ROW_KEYS = ("lig_pos", "rot_update", "tr_update", "energy", "num_clashes")


def _exchange_packed(done, plan, sizes, rank, world, group):
    """One float32 all-gather of every rank's chunk results.  Layout of a rank's buffer: its chunks in plan order, each as
    lig_pos [n, L, 3, 3] | rot_update [n, 3] | tr_update [n, 3] | energy [n] | num_clashes [n] (exact in float32 below 2^24).
    -> [(c, lo, hi, dict of CPU tensors)] for ALL chunks, on every rank."""
    import torch.distributed as dist

    def numel(c, lo, hi):
        return (hi - lo) * (int(sizes[c][1]) * 9 + 8)

    per_rank = [[ch for ch in plan if ch[3] == r] for r in range(world)]
    lens = [sum(numel(c, lo, hi) for c, lo, hi, _ in chs) for chs in per_rank]
    mine = {(c, lo, hi): d for c, lo, hi, d in done}
    dev = next((d["energy"].device for _, _, _, d in done), None)
    if dev is None:       # a rank without work still takes part in the collective, on the model's device
        dev = torch.device("cuda", torch.cuda.current_device()) if (world > 1 and dist.get_backend(group) == "nccl") else torch.device("cpu")
    flat = [t for c, lo, hi, _ in per_rank[rank] for t in
            (mine[(c, lo, hi)][k].reshape(-1).to(torch.float32) for k in _ROW_KEYS)]
    buf = torch.cat(flat) if flat else torch.zeros(0, device=dev)
    if world > 1:
        pad = torch.zeros(max(max(lens), 1), dtype=torch.float32, device=dev)
        pad[: buf.numel()] = buf
        allbuf = torch.empty(world * pad.numel(), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allbuf, pad, group=group)
        allbuf = allbuf.view(world, -1).cpu()
    else:
        allbuf = buf.cpu()[None]
    out = []
    for r in range(world):
        off = 0
        for c, lo, hi, _ in per_rank[r]:
            n, L = hi - lo, int(sizes[c][1])
            row = allbuf[r]
            d = {"lig_pos": row[off: off + n * L * 9].view(n, L, 3, 3).clone()}
            off += n * L * 9
            d["rot_update"] = row[off: off + 3 * n].view(n, 3).clone(); off += 3 * n
            d["tr_update"] = row[off: off + 3 * n].view(n, 3).clone(); off += 3 * n
            d["energy"] = row[off: off + n].clone(); off += n
            d["num_clashes"] = row[off: off + n].round().to(torch.int32); off += n
            out.append((c, lo, hi, d))
    return out


def sample_complex_set(model, loaders, sizes, num_samples, num_steps=40, eps=1e-3, use_clash_force=False,
                       noise_annealing=False, tr_noise_scale=0.5, rot_noise_scale=0.5, centre_mode=0, seed=0, ode=False,
                       group=None, min_nodes=8192, seeds=None, stats=None):
    """Many complexes x num_samples trajectories each (BASELINE config #5) over the ranks of `group`.

    loaders[c]() -> batch dict of complex c (called only on the ranks that own a chunk of it); sizes[c] = residues of
    complex c, or the pair (R, L) (for the plan, dfmdock_b200.distributed.plan_work).  Trajectory k of every complex uses Philox subsequence
    k whatever the plan, so the result does not depend on the number of ranks; seeds[c] (default: `seed` for all) is the
    Philox key of complex c -- pass distinct values so that complexes do not share initial poses and noise.  One collective at the end: the chunks'
    result rows (pose, rot_update, tr_update, energy, num_clashes) are all-gathered -- as one packed float32 tensor when sizes holds
    (R, L) pairs (the rows then stay on the device until that point), as pickled objects otherwise.
    stats: optional dict that receives this rank's wall-clock seconds per phase ("compute_s": its own chunks incl. loading the
    records and the copies back, "gather_s": the collective, "assemble_s").
    Returns (results, plan): results[c] = dict of CPU tensors {lig_pos [T,L,3,3], rot_update [T,3], tr_update [T,3],
    energy [T], num_clashes [T], best} on every rank.
    """
    rank, world = dist_utils.rank_world(group)
    plan = dist_utils.plan_work(sizes, num_samples, world, min_nodes=min_nodes)
    mine = {}
    for c, lo, hi, r in plan:
        if r == rank:
            mine.setdefault(c, []).append((lo, hi))
    import time
    t0 = time.perf_counter()
    # With (R, L) sizes every rank knows the shape of every chunk's result rows from the plan alone: the results then stay on
    # the device (no per-chunk synchronisation, so loading the next record overlaps the running chunk) and travel in ONE
    # packed float32 all-gather at the end.  With plain residue counts the rows are gathered as pickled objects.
    packed = all(isinstance(sz, (tuple, list)) for sz in sizes)
    done = []
    for c in sorted(mine):
        batch = loaders[c]()
        model.set_complex(batch)
        for lo, hi in mine[c]:
            res = model.sample(batch["lig_pos"], hi - lo, num_steps=num_steps, eps=eps, tr_noise_scale=tr_noise_scale,
                               rot_noise_scale=rot_noise_scale, use_clash_force=use_clash_force,
                               noise_annealing=noise_annealing, centre_mode=centre_mode,
                               seed=seed if seeds is None else seeds[c], stream_base=lo, ode=ode)
            if packed and tuple(res["lig_pos"].shape[:2]) != (hi - lo, int(sizes[c][1])):
                raise RuntimeError("sample_complex_set: complex %d returned poses %s, sizes says L = %d" % (c, tuple(res["lig_pos"].shape), sizes[c][1]))
            done.append((c, lo, hi, res if packed else {k: v.cpu() for k, v in res.items()}))
    t1 = time.perf_counter()
    if packed:
        everything = _exchange_packed(done, plan, sizes, rank, world, group)
    else:
        everything = [item for part in dist_utils.gather_objects(done, group) for item in part]
    t2 = time.perf_counter()
    results = []
    for c in range(len(sizes)):
        parts = sorted((lo, hi, d) for cc, lo, hi, d in everything if cc == c)
        covered = [(lo, hi) for lo, hi, _ in parts]
        if not parts or covered[0][0] != 0 or covered[-1][1] != num_samples or any(a[1] != b[0] for a, b in zip(covered, covered[1:])):
            raise RuntimeError("sample_complex_set: complex %d is not covered exactly once: %s" % (c, covered))
        out = {k: torch.cat([d[k] for _, _, d in parts], dim=0) for k in parts[0][2]}
        out["best"] = int(torch.argmin(out["energy"]).item())
        results.append(out)
    if stats is not None:
        stats.update(compute_s=t1 - t0, gather_s=t2 - t1, assemble_s=time.perf_counter() - t2)
    return results, plan
