"""Times the edge kernel at the benchmark size and reports accuracy vs the reference golden, for the variant in
DFM_EDGE_VARIANT (used to choose between kernel variants on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bench import make_workload, TRAJ_PER_GPU
from dfmdock_b200 import Score_Model
from util import FWD_CASES, load_golden, rel_err

v = os.environ.get("DFM_EDGE_VARIANT", "default")
errs = {}
for name, mk in FWD_CASES.items():
    sd, hp, batch = mk()
    model = Score_Model(sd, hp, precision="fp16").to("cuda")
    model.set_complex(batch)
    for item in load_golden(name):
        out = model.score(batch["lig_pos"][None], torch.tensor([item["t"]]), edges=item["nbr"][None].int(), want_energy=True)
        for k in ("f", "tr_score", "rot_score"):
            errs[k] = max(errs.get(k, 0), rel_err(out[k].cpu()[0], item[k].reshape(out[k].shape[1:])))
        errs["energy"] = max(errs.get("energy", 0), abs(float(out["energy"][0]) - float(item["energy"])))
real = os.path.join(ROOT, "oracle", "_ref", "golden_real.pt")
if os.path.exists(real):
    from dfmdock_b200.features import batch_from_record
    models = {}
    for g in torch.load(real, weights_only=False):
        if g["ckpt"] not in models:
            ck = torch.load(os.path.join(ROOT, "oracle", "_ref", g["ckpt"] + ".pt"), weights_only=False)
            models[g["ckpt"]] = Score_Model(ck["state_dict"], ck["hparams"], precision="fp16").to("cuda")
        model = models[g["ckpt"]]
        batch = batch_from_record(torch.load(os.path.join(ROOT, "oracle", "_ref", "db5_%s.pt" % g["complex"]), weights_only=False), pos_width=model.pos_width)
        model.set_complex(batch)
        out = model.score(batch["lig_pos"][None], torch.tensor([g["t"]]), edges=g["nbr"][None].int(), want_energy=True)
        for k in ("f", "tr_score", "rot_score"):
            errs["real_" + k] = max(errs.get("real_" + k, 0), rel_err(out[k].cpu()[0], g[k].reshape(out[k].shape[1:])))
        errs["real_energy"] = max(errs.get("real_energy", 0), abs(float(out["energy"][0]) - float(g["energy"])))
sd, hp, batch, _ = make_workload()
model = Score_Model(sd, hp, precision="fp16").to("cuda")
model.set_complex(batch)
B = TRAJ_PER_GPU
lig, tr_u, rot_u = model.randomize_pose(batch["lig_pos"], B, seed=0)
t = torch.full((B,), 0.5, device="cuda")
for i in range(2):
    model.score(lig, t, seed=0, forward_index=i)
torch.cuda.synchronize()
model.profile_enable(64)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(4):
    model.score(lig, t, seed=0, forward_index=10 + i)
e1.record()
torch.cuda.synchronize()
ms, n = model.profile_read()
print("variant %s: edge kernel %.3f ms/launch (%d launches), forward %.2f ms | worst errors %s" % (
    v, ms / n, n, e0.elapsed_time(e1) / 4, {k: "%.2e" % x for k, x in errs.items()}), flush=True)
