"""TEST INFRASTRUCTURE ONLY.  Extracts what the real-checkpoint parity tests need from the mounted reference tree into
the git-ignored oracle/_ref/ (which travels to the GPU box with the repo snapshot):
    oracle/_ref/pinder_0.pt, dips_model_0.pt   {"state_dict", "hparams"} of weights/pinder_0.ckpt, checkpoints/dips/model_0.ckpt
    oracle/_ref/db5_<id>.pt                     {"receptor": {x,pos,seq}, "ligand": {...}} of data/db5_test/<id>.pt
    oracle/_ref/golden_real.pt                  reference outputs on those inputs (live reference, injected edges)
No reference SOURCE is copied -- only data files re-serialised without the omegaconf / torch_geometric pickles.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "oracle", "_ref")
COMPLEXES = ["1QA9", "7CEI", "4POU"]


def main(quiet=False, force=False):
    from oracle import ref_shims
    if not ref_shims.reference_available():
        if not quiet:
            print("reference tree not mounted; nothing to do")
        return
    os.makedirs(OUT, exist_ok=True)
    done = os.path.join(OUT, "golden_real.pt")
    if os.path.exists(done) and not force:
        return
    from dfmdock_b200.checkpoint import load_checkpoint, load_db5_record
    from dfmdock_b200.features import batch_from_record
    root = ref_shims.REFERENCE_ROOT
    ck = {"pinder_0": os.path.join(root, "weights", "pinder_0.ckpt"), "dips_model_0": os.path.join(root, "checkpoints", "dips", "model_0.ckpt")}
    for name, path in ck.items():
        sd, hp = load_checkpoint(path)
        torch.save({"state_dict": sd, "hparams": hp}, os.path.join(OUT, name + ".pt"))
    recs = {}
    for cid in COMPLEXES:
        recs[cid] = load_db5_record(os.path.join(root, "data", "db5_test", cid + ".pt"))
        torch.save(recs[cid], os.path.join(OUT, "db5_%s.pt" % cid))
    # goldens from the live reference: real checkpoints x real complexes, graph captured and stored.  Two kinds of pose:
    #   "native": the bound pose of the record at t = 0.9 / 0.1
    #   "far":    a randomize_pose start (inference_base.py:318-340: random rotation, translation ~ N(0, 30^2) per axis,
    #             rescaled to a centroid separation drawn from 50..150 A), t = 1.0 -- what every trajectory begins from;
    #             radial = |x_i - x_j|^2 reaches 1e4..1e5 A^2 there
    # Seeds are fixed integers so that the file is reproducible.
    golden = []
    ref_shims.install()
    import models.score_net_mlsb as snm
    from oracle import dfmdock_oracle as orc
    from scipy.spatial.transform import Rotation
    for ci, (ck_name, path) in enumerate(ck.items()):
        model, hp = ref_shims.build_reference_model(path)
        width = hp.model["positional_embed_dim"]
        for xi, cid in enumerate(COMPLEXES):
            base = batch_from_record(recs[cid], pos_width=width)
            cases = [("native", 0.9, None), ("native", 0.1, None), ("far", 1.0, 60.0), ("far", 1.0, 140.0)]
            for ki, (pose, t, sep) in enumerate(cases):
                seed = 1000 + 100 * ci + 10 * xi + ki
                batch = dict(base)
                batch["t"] = torch.tensor([t])
                if pose == "far":
                    g = torch.Generator().manual_seed(seed)
                    rot0 = torch.from_numpy(Rotation.random(random_state=seed).as_matrix()).float()
                    tr0 = torch.randn(1, 3, generator=g)
                    tr0 = tr0 / tr0.norm() * sep          # separation of the CA centroids after randomize_pose
                    batch["lig_pos"], _, _ = orc.randomize_pose(base["rec_pos"], base["lig_pos"], rot0, tr0)
                captured = {}
                orig = snm.get_knn_and_sample

                def gk(points, *a, **k):
                    out = orig(points, *a, **k)
                    captured["nbr"] = torch.cat([o for o in out if o is not None], dim=-1).clone()
                    return out

                snm.get_knn_and_sample = gk
                try:
                    torch.manual_seed(seed)
                    with torch.no_grad():
                        out = model(batch)
                finally:
                    snm.get_knn_and_sample = orig
                golden.append({"ckpt": ck_name, "complex": cid, "t": t, "pose": pose, "sep": sep, "seed": seed,
                               "lig_pos": batch["lig_pos"].clone(), "nbr": captured["nbr"].to(torch.int16),
                               "tr_score": out["tr_score"], "rot_score": out["rot_score"], "energy": out["energy"],
                               "f": out["f"], "num_clashes": out["num_clashes"]})
    torch.save(golden, done)
    if not quiet:
        print("wrote", OUT, [(g["ckpt"], g["complex"], g["pose"], g["t"], float(g["energy"])) for g in golden])


def extract_all_db5(quiet=False):
    """BASELINE config #5 inputs: every complex of data/db5_test/test.txt -> oracle/_ref/db5_all/<id>.pt (65 MB, git-ignored;
    only needed for profiles/run_db5_set.py -- remove the directory afterwards to keep gpurun snapshots small)."""
    from oracle import ref_shims
    if not ref_shims.reference_available():
        print("reference tree not mounted; nothing to do")
        return
    from dfmdock_b200.checkpoint import load_db5_record
    root = os.path.join(ref_shims.REFERENCE_ROOT, "data", "db5_test")
    out = os.path.join(OUT, "db5_all")
    os.makedirs(out, exist_ok=True)
    ids = open(os.path.join(root, "test.txt")).read().split()
    for cid in ids:
        rec = load_db5_record(os.path.join(root, cid + ".pt"))
        torch.save(rec, os.path.join(out, cid + ".pt"))
        if not quiet:
            print(cid, rec["receptor"]["pos"].shape[0], rec["ligand"]["pos"].shape[0])


if __name__ == "__main__":
    if "--all-db5" in sys.argv:
        extract_all_db5()
    else:
        main(force="--force" in sys.argv)
