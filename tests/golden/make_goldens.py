"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference, via oracle/ref_shims.py).

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_goldens.py
Committed outputs are small: they hold the reference's outputs plus the noise / neighbour tables that were
injected; the inputs are regenerated at test time from seeds (dfmdock_b200.synthetic / features) or, for the
real-checkpoint golden, read from the git-ignored oracle/_ref/ (oracle/build_ref.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from dfmdock_b200.features import synthetic_complex  # noqa: E402
from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(8)


def build_ref_model(sd, hp):
    ref_shims.install()
    from models.score_model_mlsb import Score_Model

    hp = ref_shims.AttrDict(hp)
    hp["diffuser"]["so3"]["cache_dir"] = "/tmp/dfmdock_so3_cache"
    m = Score_Model(hp.model, hp.diffuser, hp.experiment)
    m.load_state_dict({"net." + k: v for k, v in sd.items()}, strict=True)
    return m.eval()


class Recorder:
    """Monkey-patches the RNG entry points the hot path uses and records every draw."""

    def __init__(self, tr_std=None):
        self.exp, self.randn, self.normal, self.rot0, self.nbr = [], [], [], [], []
        self.tr_std = tr_std      # overrides the N(0, 30^2) initial translation so that the chains start in contact

    def __enter__(self):
        import models.score_net_mlsb as snm
        from scipy.spatial.transform import Rotation

        self.snm, self.Rotation = snm, Rotation
        self._mn, self._randn, self._normal = torch.multinomial, torch.randn, torch.normal
        self._rr, self._gk = Rotation.random, snm.get_knn_and_sample
        rec = self

        def multinomial(p, k, replacement=False):
            q = torch.empty_like(p).exponential_(1)      # identical draw to torch.multinomial's no-replacement path
            rec.exp.append(q.clone())
            return torch.topk(p / q, k).indices

        def randn(*a, **k):
            z = rec._randn(*a, **k)
            if tuple(z.shape) == (1, 3):
                rec.randn.append(z.clone())
            return z

        def normal(*a, **k):
            if rec.tr_std is not None:
                a = (a[0], rec.tr_std) + tuple(a[2:])
            z = rec._normal(*a, **k)
            rec.normal.append(z.clone())
            return z

        def rot_random(*a, **k):
            r = rec._rr(*a, **k)
            rec.rot0.append(torch.from_numpy(r.as_matrix()).float())
            return r

        def gk(points, *a, **k):
            out = rec._gk(points, *a, **k)
            rec.nbr.append(torch.cat([o for o in out if o is not None], dim=-1).clone())
            return out

        torch.multinomial, torch.randn, torch.normal = multinomial, randn, normal
        Rotation.random = staticmethod(rot_random)
        snm.get_knn_and_sample = gk
        return self

    def __exit__(self, *exc):
        torch.multinomial, torch.randn, torch.normal = self._mn, self._randn, self._normal
        self.Rotation.random = self._rr
        self.snm.get_knn_and_sample = self._gk


def check_multinomial_identity():
    p = torch.rand(50, 300) + 1e-3
    torch.manual_seed(7)
    a = torch.multinomial(p, 40, replacement=False)
    torch.manual_seed(7)
    q = torch.empty_like(p).exponential_(1)
    b = torch.topk(p / q, 40).indices
    assert torch.equal(a, b), "torch.multinomial is no longer topk(p / Exp(1)) on this torch build"


def layer_hooks(model, store):
    hs = []
    for i in range(6):
        layer = model.net.network._modules["EGNN_%d" % i]
        hs.append(layer.register_forward_hook(lambda mod, inp, out, i=i: store.__setitem__("h%d" % i, out[0].detach().clone())))
    return hs


def forward_golden(name, sd, hp, batch, t_values, seed):
    model = build_ref_model(sd, hp)
    import models.score_net_mlsb as snm

    items = []
    for t in t_values:
        b = dict(batch)
        b["t"] = torch.tensor([t])
        store = {}
        hooks = layer_hooks(model, store)
        torch.manual_seed(seed)
        with Recorder() as rec, torch.no_grad():
            out = model(b)
        for h in hooks:
            h.remove()
        pos = torch.cat([b["rec_pos"], b["lig_pos"]], dim=0)
        centred = pos - b["lig_pos"][:, 1, :].mean(dim=0)
        sm = snm.get_spatial_matrix(centred)
        bins = torch.stack([sm[..., 0:40].argmax(-1), sm[..., 40:64].argmax(-1), sm[..., 64:88].argmax(-1),
                            sm[..., 88:100].argmax(-1)], dim=0).to(torch.uint8)
        items.append({
            "t": t, "exp": rec.exp[0].half().float() if False else rec.exp[0], "nbr": rec.nbr[0].to(torch.int16),
            "tr_score": out["tr_score"], "rot_score": out["rot_score"], "energy": out["energy"], "f": out["f"],
            "num_clashes": out["num_clashes"], "bins": bins,
            "h": torch.stack([store["h%d" % i] for i in range(6)], dim=0).half(),
        })
        seed += 1
    torch.save(items, os.path.join(OUT, name))
    print(name, "written:", [(it["t"], float(it["energy"])) for it in items])


def sampler_golden(name, sd, hp, batch, num_steps, seed, use_clash_force, variant, tr_std=None):
    model = build_ref_model(sd, hp)
    if variant == "inference_base":
        import inference_base as mod
    else:
        import inference as mod
    import random

    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    steps = []
    orig_modify = mod.modify_coords

    def modify(x, rot, tr):
        y = orig_modify(x, rot, tr)
        steps.append({"rot": rot.clone(), "tr": tr.clone()})
        return y

    scores = []
    orig_fwd = model.forward

    def fwd(b):
        o = orig_fwd(b)
        scores.append({"tr_score": o["tr_score"].clone(), "rot_score": o["rot_score"].clone(), "energy": o["energy"].clone(),
                       "lig_pos": b["lig_pos"].clone(), "num_clashes": o["num_clashes"].clone()})
        return o

    model.forward = fwd
    mod.modify_coords = modify
    try:
        with Recorder(tr_std) as rec:
            rec_pos, lig_pos, rot_update, tr_update, out = mod.Euler_Maruyama_sampler(
                model, dict(batch), num_steps=num_steps, device="cpu", use_clash_force=use_clash_force)
    finally:
        mod.modify_coords = orig_modify
    g = {
        "variant": variant, "num_steps": num_steps, "use_clash_force": use_clash_force,
        "rot0": rec.rot0[0], "tr0": rec.normal[0], "exp": torch.stack(rec.exp, 0), "nbr": torch.stack(rec.nbr, 0).to(torch.int16),
        "z": torch.stack(rec.randn, 0).view(num_steps, 2, 3),
        "fwd_lig_pos": torch.stack([s["lig_pos"] for s in scores], 0),
        "tr_score": torch.stack([s["tr_score"] for s in scores], 0), "rot_score": torch.stack([s["rot_score"] for s in scores], 0),
        "step_rot": torch.stack([s["rot"] for s in steps], 0), "step_tr": torch.stack([s["tr"] for s in steps], 0),
        "tr_std": tr_std, "lig_pos": lig_pos, "rot_update": rot_update, "tr_update": tr_update, "energy": out["energy"],
        "num_clashes": out["num_clashes"],
    }
    torch.save(g, os.path.join(OUT, name))
    print(name, "written: energy", float(out["energy"]), "rot_update", rot_update.tolist())


def clash_golden(name, batch):
    ref_shims.install()
    import inference_base as ib
    import inference as inf

    items = []
    for shift in (4.0, 9.0, 60.0):
        lig = batch["lig_pos"] - batch["lig_pos"][:, 1].mean(0) + batch["rec_pos"][:, 1].mean(0) + torch.tensor([shift, 0.0, 0.0])
        f0 = ib.get_clash_force(batch["rec_pos"].clone(), lig.clone())
        f1 = inf.get_clash_force(batch["rec_pos"].clone(), lig.clone())
        assert torch.equal(f0, f1)
        d = (batch["rec_pos"].reshape(-1, 1, 3) - lig.reshape(1, -1, 3)).norm(dim=-1)
        items.append({"shift": shift, "lig_pos": lig, "force": f0, "pairs_lt4": int((d < 4).sum())})
    torch.save(items, os.path.join(OUT, name))
    print(name, "written:", [(it["shift"], it["pairs_lt4"], it["force"].tolist()) for it in items])


def main():
    check_multinomial_identity()
    hp66, hp67 = synthetic_hparams(66), synthetic_hparams(67)
    sd66, sd67 = synthetic_state_dict(1, 66), synthetic_state_dict(2, 67)
    small = synthetic_complex(40, 30, seed=3, pos_width=66)
    # bring the chains into contact so that the energy mask / 6D bins / clashes are exercised
    small["lig_pos"] = small["lig_pos"] - torch.tensor([17.0, 0.0, 0.0])
    forward_golden("fwd_synth_n70.pt", sd66, hp66, small, [0.8, 0.05], seed=11)
    mid = synthetic_complex(75, 53, seed=4, pos_width=67)
    mid["lig_pos"] = mid["lig_pos"] - torch.tensor([15.0, 0.0, 0.0])
    forward_golden("fwd_synth_n128_p67.pt", sd67, hp67, mid, [0.5], seed=21)
    tiny = synthetic_complex(25, 20, seed=5, pos_width=66)          # 20 <= N < 60: 25 sampled edges
    tiny["lig_pos"] = tiny["lig_pos"] - torch.tensor([18.0, 0.0, 0.0])
    forward_golden("fwd_synth_n45.pt", sd66, hp66, tiny, [0.3], seed=31)
    clash_golden("clash_force_n70.pt", small)
    sampler_golden("sampler_base_n70.pt", sd66, hp66, small, 5, seed=42, use_clash_force=False, variant="inference_base")
    sampler_golden("sampler_clash_n70.pt", sd66, hp66, small, 4, seed=43, use_clash_force=True, variant="inference", tr_std=3.0)
    sampler_golden("sampler_near_n70.pt", sd66, hp66, small, 4, seed=44, use_clash_force=False, variant="inference_base", tr_std=6.0)


if __name__ == "__main__":
    main()
