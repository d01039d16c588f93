import os, sys, time, torch
sys.path.insert(0, ".")
from bench import make_workload, TRAJ_PER_GPU, time_full_job
from dfmdock_b200 import Score_Model
sd, hp, batch, desc = make_workload()
model = Score_Model(sd, hp, precision="fp16").to("cuda")
model.set_complex(batch)
for S in (100, 100, 50, 10, 2):
    ms, res = time_full_job(model, batch, TRAJ_PER_GPU, S, 0)
    print("steps %3d: %.1f ms -> %.3f ms per forward (S+1 forwards)" % (S, ms, ms / (S + 1)))
# per-step time along a trajectory: far start vs contact
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.sample(batch["lig_pos"], TRAJ_PER_GPU, num_steps=100, seed=0)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "k_prepare" in e.name], key=lambda e: e.time_range.start)
st = [e.time_range.start for e in evs]
d = [(b - a) / 1e3 for a, b in zip(st[:-1], st[1:])]
print("step durations ms (k_prepare to k_prepare): first 5", ["%.2f" % x for x in d[:5]], "mid", ["%.2f" % x for x in d[48:52]], "last 5", ["%.2f" % x for x in d[-5:]])
import collections
agg = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA: agg[e.name[:50]] += e.device_time_total
for k, v in agg.most_common(8): print("%-52s %10.1f us" % (k, v))
