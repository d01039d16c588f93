"""Short driver for ncu: a few lock-step steps of the benchmark workload (bench.py's timed state: 2x150 residues, 256
trajectories in contact, pinder_0 weights when oracle/_ref holds them).  PROFILE_CONFIG=c2 runs BASELINE config #2 instead
(db5 1QA9, 40 trajectories, clash force) through dfm_sample."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import REF_DIR, TRAJ_PER_GPU, contact_poses, make_workload
from dfmdock_b200 import Score_Model

sd, hp, batch, desc = make_workload()
model = Score_Model(sd, hp, precision="fp16").to("cuda")
n_fwd = int(os.environ.get("PROFILE_FORWARDS", 2))
if os.environ.get("PROFILE_CONFIG") == "c2":
    from dfmdock_b200.features import batch_from_record
    batch = batch_from_record(torch.load(os.path.join(REF_DIR, "db5_1QA9.pt"), weights_only=False), pos_width=model.pos_width)
    model.set_complex(batch)
    model.sample(batch["lig_pos"], 40, num_steps=max(2, n_fwd), seed=1, use_clash_force=True, centre_mode=1)
else:
    model.set_complex(batch)
    B = int(os.environ.get("PROFILE_B", TRAJ_PER_GPU))
    lig = contact_poses(batch["lig_pos"], B, seed=1000).cuda()
    tr_u, rot_u = torch.zeros(B, 3, device="cuda"), torch.zeros(B, 3, device="cuda")
    t = torch.full((B,), 0.3, device="cuda")
    for i in range(n_fwd):
        o = model.score(lig, t, seed=0, forward_index=i, want_energy=bool(os.environ.get("PROFILE_ENERGY")))
        model.reverse_step(lig, rot_u, tr_u, o["tr_score"], o["rot_score"], 0.3, 0.01, 0.5, 0.5, seed=0, step_index=i)
torch.cuda.synchronize()
print("done", desc)
