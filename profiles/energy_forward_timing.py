import os, sys, torch
sys.path.insert(0, ".")
from bench import make_workload, contact_poses, TRAJ_PER_GPU
from dfmdock_b200 import Score_Model
sd, hp, batch, desc = make_workload()
model = Score_Model(sd, hp, precision="fp16").to("cuda")
model.set_complex(batch)
B = TRAJ_PER_GPU
lig = contact_poses(batch["lig_pos"], B, seed=1000).cuda()
t = torch.full((B,), 0.3, device="cuda")
def run(we, n=5):
    for i in range(2): model.score(lig, t, seed=0, forward_index=i, want_energy=we)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): model.score(lig, t, seed=0, forward_index=i, want_energy=we)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("forward without energy %.3f ms, with energy %.3f ms" % (run(False), run(True)))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.score(lig, t, seed=0, forward_index=9, want_energy=True)
    torch.cuda.synchronize()
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]:
    print("%-60s n=%2d %9.1f us" % (ev.key[:60], ev.count, ev.device_time_total))
