"""Per-kernel durations inside real lock-step steps of BASELINE config #3 (2x150 residues, 256 trajectories) from CUPTI activity
records (torch.profiler): warm caches, kernels back to back -- unlike the ncu launch list, whose launches are serialised with
cold caches.  Run with DFM_PDL=0 for durations that are not stretched by programmatic-dependent-launch overlap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from bench import make_workload, TRAJ_PER_GPU
from dfmdock_b200 import Score_Model

sd, hp, batch, _ = make_workload()
model = Score_Model(sd, hp, precision="fp16").to("cuda")
model.set_complex(batch)
B = int(os.environ.get("TRAJ", TRAJ_PER_GPU))
model.sample(batch["lig_pos"], B, num_steps=4, seed=1)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.sample(batch["lig_pos"], B, num_steps=8, seed=2)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0 and "mem" not in e.name.lower()],
             key=lambda e: e.time_range.start)
starts = [i for i, e in enumerate(evs) if "k_prepare" in e.name]
a, b = starts[3], starts[5]
agg = {}
for e in evs[a:b]:
    n = e.name.split("(")[0][-40:]
    d = agg.setdefault(n, [0, 0.0])
    d[0] += 1; d[1] += e.time_range.end - e.time_range.start
span = evs[b].time_range.start - evs[a].time_range.start
print("PDL=%s B=%d: two steps span %.1f us (%.1f per step)" % (os.environ.get("DFM_PDL", "1"), B, span, span / 2))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-42s n=%3d  %8.1f us per step  %7.1f us per launch  %5.1f %%" % (n, c // 2, t / 2, t / c, 100 * t / span))
