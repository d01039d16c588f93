"""Launch-time breakdown on the largest db5 complex (1N2C, N = 2548, 40 trajectories): run under
ncu --metrics gpu__time_duration.sum to see which kernels dominate outside the benchmark size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dfmdock_b200 import Score_Model
from dfmdock_b200.features import batch_from_record

ref = os.path.join(ROOT, "oracle", "_ref")
ck = torch.load(os.path.join(ref, "pinder_0.pt"), weights_only=False)
model = Score_Model(ck["state_dict"], ck["hparams"], precision="fp16").to("cuda")
cid = os.environ.get("PROFILE_COMPLEX", "1N2C")
batch = batch_from_record(torch.load(os.path.join(ref, "db5_%s.pt" % cid), weights_only=False), pos_width=model.pos_width)
model.set_complex(batch)
B = int(os.environ.get("PROFILE_B", 40))
lig, tr_u, rot_u = model.randomize_pose(batch["lig_pos"], B, seed=0, centre_mode=1)
t = torch.full((B,), 0.5, device="cuda")
for i in range(2):
    o = model.score(lig, t, seed=0, forward_index=i, want_energy=(i == 1))
    model.reverse_step(lig, rot_u, tr_u, o["tr_score"], o["rot_score"], 0.5, 0.01, 0.5, 0.5, seed=0, step_index=i,
                       use_clash_force=True, centre_mode=1)
torch.cuda.synchronize()
print("done", cid, B)
