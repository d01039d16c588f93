"""Short driver for ncu: a few score-network forwards at the benchmark size (2x150, 256 trajectories)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_workload, TRAJ_PER_GPU
from dfmdock_b200 import Score_Model

sd, hp, batch = make_workload()
model = Score_Model(sd, hp, precision="fp16").to("cuda")
model.set_complex(batch)
B = int(os.environ.get("PROFILE_B", TRAJ_PER_GPU))
lig, tr_u, rot_u = model.randomize_pose(batch["lig_pos"], B, seed=0)
t = torch.full((B,), 0.5, device="cuda")
for i in range(int(os.environ.get("PROFILE_FORWARDS", 2))):
    o = model.score(lig, t, seed=0, forward_index=i)
    model.reverse_step(lig, rot_u, tr_u, o["tr_score"], o["rot_score"], 0.5, 0.01, 0.5, 0.5, seed=0, step_index=i)
torch.cuda.synchronize()
print("done")
