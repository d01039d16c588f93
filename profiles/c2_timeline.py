"""Timeline of BASELINE config #2 (db5 1QA9, N = 197, 40 trajectories, clash force) from CUPTI activity records (torch.profiler;
nsys is not in the image): every kernel of two consecutive lock-step steps of dfm_sample with start offset, duration and the
idle gap before it, then the totals (busy time, gaps, launches per step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from dfmdock_b200 import Score_Model
from dfmdock_b200.features import batch_from_record

ref = os.path.join(ROOT, "oracle", "_ref")
ck = torch.load(os.path.join(ref, "pinder_0.pt"), weights_only=False)
model = Score_Model(ck["state_dict"], ck["hparams"], precision="fp16").to("cuda")
batch = batch_from_record(torch.load(os.path.join(ref, "db5_1QA9.pt"), weights_only=False), pos_width=model.pos_width)
model.set_complex(batch)
kw = dict(use_clash_force=True, centre_mode=1)
model.sample(batch["lig_pos"], 40, num_steps=5, seed=1, **kw)
torch.cuda.synchronize()
STEPS = 12
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.sample(batch["lig_pos"], 40, num_steps=STEPS, seed=2, **kw)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0 and "mem" not in e.name.lower()],
             key=lambda e: e.time_range.start)
# one lock-step step = from one k_prepare to the next
starts = [i for i, e in enumerate(evs) if "k_prepare" in e.name]
print("kernels recorded %d, steps found %d" % (len(evs), len(starts)))
if len(starts) >= 6:
    a, b = starts[3], starts[5]
    t0 = evs[a].time_range.start
    prev_end = t0
    busy = gaps = 0.0
    for e in evs[a:b]:
        st, du = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = e.time_range.start - prev_end
        print("%9.1f us  dur %7.1f us  gap %6.1f us  %s" % (st, du, gap, e.name[:70]))
        busy += du; gaps += max(gap, 0.0); prev_end = max(prev_end, e.time_range.end)
    span = evs[b].time_range.start - t0
    print("two steps: span %.1f us (%.1f us per step), kernel time %.1f us, idle gaps %.1f us (negative gaps = overlap under programmatic dependent launch), %d launches per step"
          % (span, span / 2, busy, gaps, (b - a) // 2))
