"""TEST INFRASTRUCTURE.  T4 of SURVEY.md 8(c): a FREE-RUNNING 10-step trajectory of the UNMODIFIED reference sampler with
every random draw recorded, so that the CUDA sampler can be run on the same noise (injected neighbour tables + z) and
its final pose compared with the reference's (CA-RMSD <= 0.05 A, final energy).

    python tests/golden/make_t4_golden.py

Build container only (needs /root/reference through oracle/ref_shims.py).  Inputs are the real checkpoints and the
1QA9 record that oracle/build_ref.py extracts to oracle/_ref (they travel to the GPU box); outputs are small:
    tests/golden/t4_1QA9_dips_s10.pt            src/inference_base.py sampler, checkpoints/dips/model_0.ckpt
    tests/golden/t4_1QA9_pinder_s10_clash.pt    src/inference.py sampler (all-atom centroid, clash force), weights/pinder_0.ckpt
each = {"rot0", "tr0", "nbr" [S+1,N,60] int16, "z" [S,2,3], "fwd_lig_pos" [S+1,L,3,3], "tr_score", "rot_score",
        "lig_pos", "rot_update", "tr_update", "energy", "num_clashes", "num_steps", "variant", "use_clash_force", "ckpt"}.
The initial translation is drawn with std 8 A instead of 30 A so that the chains interact from the first step (the far
regime is covered by the "far" cases of oracle/_ref/golden_real.pt).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_goldens import Recorder  # noqa: E402
from oracle import ref_shims  # noqa: E402


def make(name, ckpt_rel, variant, use_clash_force, seed, num_steps=10, tr_std=8.0):
    from dfmdock_b200.features import batch_from_record
    ref_shims.install()
    if variant == "inference_base":
        import inference_base as mod
    else:
        import inference as mod
    root = ref_shims.REFERENCE_ROOT
    model, hp = ref_shims.build_reference_model(os.path.join(root, ckpt_rel))
    rec = ref_shims.load_db5_record(os.path.join(root, "data", "db5_test", "1QA9.pt"))
    batch = batch_from_record(rec, pos_width=hp.model["positional_embed_dim"])
    import random
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    scores = []
    orig_fwd = model.forward

    def fwd(b):
        o = orig_fwd(b)
        scores.append({"tr_score": o["tr_score"].clone(), "rot_score": o["rot_score"].clone(), "lig_pos": b["lig_pos"].clone()})
        return o

    model.forward = fwd
    mod.tqdm = lambda x, **k: x
    with Recorder(tr_std) as rec_, torch.no_grad():
        rec_pos, lig_pos, rot_update, tr_update, out = mod.Euler_Maruyama_sampler(
            model, dict(batch), num_steps=num_steps, device="cpu", use_clash_force=use_clash_force)
    g = {
        "variant": variant, "num_steps": num_steps, "use_clash_force": use_clash_force, "ckpt": os.path.basename(ckpt_rel),
        "rot0": rec_.rot0[0], "tr0": rec_.normal[0], "nbr": torch.stack(rec_.nbr, 0).to(torch.int16),
        "z": torch.stack(rec_.randn, 0).view(num_steps, 2, 3),
        "fwd_lig_pos": torch.stack([s["lig_pos"] for s in scores], 0),
        "tr_score": torch.stack([s["tr_score"] for s in scores], 0), "rot_score": torch.stack([s["rot_score"] for s in scores], 0),
        "lig_pos": lig_pos, "rot_update": rot_update, "tr_update": tr_update, "energy": out["energy"],
        "num_clashes": out["num_clashes"], "tr_std": tr_std, "seed": seed,
    }
    torch.save(g, os.path.join(HERE, name))
    native = batch["lig_pos"][:, 1]
    print(name, "energy %.4f" % float(out["energy"]), "clashes", int(out["num_clashes"]),
          "L-RMSD to native %.2f" % float(((lig_pos[:, 1] - native) ** 2).sum(-1).mean().sqrt()),
          "|tr_update| %.2f" % float(tr_update.norm()))


if __name__ == "__main__":
    torch.set_num_threads(8)
    make("t4_1QA9_dips_s10.pt", os.path.join("checkpoints", "dips", "model_0.ckpt"), "inference_base", False, seed=101)
    make("t4_1QA9_pinder_s10_clash.pt", os.path.join("weights", "pinder_0.ckpt"), "inference", True, seed=102)
