import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from dfmdock_b200 import Score_Model
from dfmdock_b200.features import synthetic_complex
from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
sd, hp = synthetic_state_dict(0, 66), synthetic_hparams(66)
for n_rec, n_lig in ((300, 260), (600, 300), (500, 100), (1000, 24)):
    batch = synthetic_complex(n_rec, n_lig, seed=2)
    batch["lig_pos"] = batch["lig_pos"] - torch.tensor([12.0, 0.0, 0.0])
    model = Score_Model(sd, hp, precision="fp16").to("cuda")
    model.set_complex(batch)
    lig = batch["lig_pos"][None]
    t = torch.full((1,), 0.4)
    a = model.score(lig, t, seed=3, stream_base=7, forward_index=5, return_edges=True)["edges"][0].cpu().long()
    model.graph_generic = True
    b = model.score(lig, t, seed=3, stream_base=7, forward_index=5, return_edges=True)["edges"][0].cpu().long()
    sa, sb = a[:, 20:].sort(-1).values, b[:, 20:].sort(-1).values
    bad = (sa != sb).any(dim=1).nonzero().flatten()
    N = n_rec + n_lig
    pos = torch.cat([batch["rec_pos"], batch["lig_pos"]], 0)[:, 1].double()
    pos = pos  # centred differently on device, distances identical
    print("N", N, "rows differing", len(bad), "of", N)
    for r in bad[:6].tolist():
        only_a = sorted(set(sa[r].tolist()) - set(sb[r].tolist())); only_b = sorted(set(sb[r].tolist()) - set(sa[r].tolist()))
        da = [(j, round(float((pos[r] - pos[j]).norm()), 3)) for j in only_a]; db = [(j, round(float((pos[r] - pos[j]).norm()), 3)) for j in only_b]
        print("  row", r, "only sel:", da, "only generic:", db)
