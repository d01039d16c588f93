"""BASELINE config #5: the full data/db5_test set (25 complexes, N = 197 .. 2548), 40 trajectories x 40 steps each with
weights/pinder_0.ckpt + clash force, through the public API (sample_trajectories + compute_metrics_batch).

    python profiles/run_db5_set.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 profiles/run_db5_set.py   # trajectory-sharded

Inputs: oracle/_ref/pinder_0.pt and oracle/_ref/db5_all/<id>.pt (python oracle/build_ref.py --all-db5 in the build
container; falls back to the three complexes of oracle/_ref/).  Prints one line per complex and the set totals; writes
gpurun_out/db5_c5.csv with the reference's CSV columns for the lowest-energy sample of each complex.
"""
import csv
import glob
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dfmdock_b200 import Score_Model  # noqa: E402
from dfmdock_b200.features import batch_from_record  # noqa: E402
from dfmdock_b200.inference import init_distributed  # noqa: E402
from dfmdock_b200.metrics import KEYS, compute_metrics_batch  # noqa: E402
from dfmdock_b200.sampler import sample_complex_set  # noqa: E402

T, S = int(os.environ.get("C5_SAMPLES", "40")), int(os.environ.get("C5_STEPS", "40"))


def main():
    world = init_distributed()
    rank = torch.distributed.get_rank() if world > 1 else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    ref = os.path.join(ROOT, "oracle", "_ref")
    ck = torch.load(os.path.join(ref, "pinder_0.pt"), weights_only=False)
    model = Score_Model(ck["state_dict"], ck["hparams"], precision="fp16").to(dev)
    paths = sorted(glob.glob(os.path.join(ref, "db5_all", "*.pt"))) or sorted(glob.glob(os.path.join(ref, "db5_*.pt")))
    ids = [os.path.splitext(os.path.basename(p))[0].replace("db5_", "") for p in paths]
    recs = [torch.load(p, weights_only=False) for p in paths]
    sizes = [(r["receptor"]["pos"].shape[0], r["ligand"]["pos"].shape[0]) for r in recs]      # (R, L) for the planner
    loaders = [(lambda r=r: batch_from_record(r, pos_width=model.pos_width, with_position_matrix=False)) for r in recs]
    # warm-up (allocator, module load) on the smallest complex
    b0 = loaders[min(range(len(sizes)), key=lambda c: sum(sizes[c]))]()
    model.set_complex(b0)
    model.sample(b0["lig_pos"], 2, num_steps=3, use_clash_force=True, centre_mode=1, seed=1)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stats = {}
    results, plan = sample_complex_set(model, loaders, sizes, T, num_steps=S, use_clash_force=True, centre_mode=1, seed=42, stats=stats)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    mine = sum(1 for ch in plan if ch[3] == rank)
    # device-side busy time of this rank's own chunks is not separable from the final object gather, so report both the
    # whole call (max over ranks) and the wall clock
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(ms)
    from dfmdock_b200.distributed import gather_objects
    all_stats = gather_objects({k: round(v * 1e3, 1) for k, v in stats.items()})
    if rank == 0:
        print("per-rank phases (ms):", all_stats)
        rows = []
        for c, res in enumerate(results):
            rec = recs[c]
            rec_pos, lig0 = rec["receptor"]["pos"].float(), rec["ligand"]["pos"].float()
            met = compute_metrics_batch(rec_pos, res["lig_pos"], rec_pos, lig0, device=dev).cpu()
            best = res["best"]
            row = {"id": ids[c], "index": str(best), "n_rec": rec_pos.shape[0], "n_lig": lig0.shape[0]}
            row.update({k: float(met[best, j]) for j, k in enumerate(KEYS)})
            row.update({"energy": float(res["energy"][best]), "num_clashes": int(res["num_clashes"][best]),
                        "best_dockq_of_40": float(torch.nan_to_num(met[:, 4], nan=0.0).max()),
                        "chunks": " ".join("%d-%d@%d" % (lo, hi, r) for cc, lo, hi, r in plan if cc == c),
                        "energy_checksum": "%.6f" % float(res["energy"].double().sum())})
            rows.append(row)
            print("%-5s R=%4d L=%4d  lowest energy %8.2f (sample %2d: DockQ %.3f, L-RMSD %6.2f)  best DockQ of %d: %.3f  chunks %s"
                  % (ids[c], row["n_rec"], row["n_lig"], row["energy"], best, row["DockQ"], row["l_rmsd"], T, row["best_dockq_of_40"],
                     row["chunks"]), flush=True)
        total_ps = len(rows) * T * S
        print("c5 total: %d complexes x %d traj x %d steps on %d GPU(s), %d chunks (%d on rank 0): %.2f s device time (max over ranks, "
              "incl. H2D of the records and the final gather), wall %.2f s; %.0f pose-steps/s"
              % (len(rows), T, S, world, len(plan), mine, ms * 1e-3, wall, total_ps / (ms * 1e-3)))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "db5_c5_%dgpu.csv" % world), "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
            w.writeheader()
            w.writerows(rows)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
