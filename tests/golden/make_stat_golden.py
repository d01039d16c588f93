"""TEST INFRASTRUCTURE.  Draws the reference sampler's own output distribution for the statistical parity test (T5 of
SURVEY.md 8c): the UNMODIFIED reference (src/inference_base.py Euler_Maruyama_sampler + Score_Model, imported through
oracle/ref_shims.py) run free on a real complex with a real checkpoint, on the build container's CPU.

    python tests/golden/make_stat_golden.py [--traj 96] [--steps 20] [--complex 1QA9] [--ckpt dips] [--variant base|inference]

--variant inference uses src/inference.py's sampler (all-atom-centroid convention) with use_clash_force=True, the way
src/inference.py:535 runs it: together with --ckpt pinder --steps 40 that is BASELINE config #2.

Writes tests/golden/stat_<complex>_<ckpt>_s<steps>.pt = {"energy"[T], "l_rmsd"[T], "tr_norm"[T], "rot_angle"[T],
"num_clashes"[T], "complex", "ckpt", "num_steps", "seed"}.  l_rmsd = CA RMSD of the final ligand pose to the pose in the
record (receptor frame fixed).  The inputs (checkpoint, record) are the ones oracle/build_ref.py extracts to oracle/_ref.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--traj", type=int, default=96)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--complex", default="1QA9")
    ap.add_argument("--ckpt", default="dips")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--variant", default="base", choices=["base", "inference"])
    args = ap.parse_args()
    from oracle import ref_shims
    from dfmdock_b200.features import batch_from_record
    ref_shims.install()
    if args.variant == "inference":
        import inference as ib       # src/inference.py
    else:
        import inference_base as ib  # src/inference_base.py
    root = ref_shims.REFERENCE_ROOT
    path = {"dips": os.path.join(root, "checkpoints", "dips", "model_0.ckpt"), "pinder": os.path.join(root, "weights", "pinder_0.ckpt")}[args.ckpt]
    model, hp = ref_shims.build_reference_model(path)
    rec = ref_shims.load_db5_record(os.path.join(root, "data", "db5_test", args.complex + ".pt"))
    batch = batch_from_record(rec, pos_width=hp.model["positional_embed_dim"])
    native = batch["lig_pos"][:, 1].double()
    ib.set_seed(args.seed)
    torch.set_num_threads(os.cpu_count() or 1)
    out = {k: [] for k in ("energy", "l_rmsd", "tr_norm", "rot_angle", "num_clashes")}
    t0 = time.time()
    ib.tqdm = lambda x, **k: x       # silence the progress bar
    for i in range(args.traj):
        rec_pos, lig_pos, rot_update, tr_update, output = ib.Euler_Maruyama_sampler(
            model=model, batch=dict(batch), num_steps=args.steps, device="cpu", use_clash_force=(args.variant == "inference"))
        out["energy"].append(float(output["energy"]))
        out["num_clashes"].append(int(output["num_clashes"]))
        out["l_rmsd"].append(float(((lig_pos[:, 1].double() - native) ** 2).sum(-1).mean().sqrt()))
        out["tr_norm"].append(float(tr_update.norm()))
        out["rot_angle"].append(float(rot_update.norm()))
        if i % 8 == 7:
            print("%d/%d trajectories, %.0f s" % (i + 1, args.traj, time.time() - t0), flush=True)
    res = {k: torch.tensor(v) for k, v in out.items()}
    res.update(complex=args.complex, ckpt=args.ckpt, num_steps=args.steps, seed=args.seed, variant=args.variant)
    tag = "" if args.variant == "base" else "_clash"
    name = os.path.join(HERE, "stat_%s_%s_s%d%s.pt" % (args.complex, args.ckpt, args.steps, tag))
    torch.save(res, name)
    print("written", name, {k: (float(v.float().mean()), float(v.float().std())) for k, v in res.items() if torch.is_tensor(v)})


if __name__ == "__main__":
    main()
