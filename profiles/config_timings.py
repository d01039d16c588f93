"""Timings of the other BASELINE configs (parity-test sizes, not bench lines): full dfm_sample jobs with CUDA events.
c2: real 1QA9 (N=197) + pinder_0 ckpt, 40 x 40, clash force (needs oracle/_ref);  c4: synthetic 2x400, 64 x 100."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dfmdock_b200 import Score_Model
from dfmdock_b200.features import batch_from_record, synthetic_complex
from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict


def timed(model, batch, T, S, **kw):
    model.set_complex(batch)
    model.sample(batch["lig_pos"], T, num_steps=3, seed=1, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = model.sample(batch["lig_pos"], T, num_steps=S, seed=2, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return ms, T * S / (ms * 1e-3), float(res["energy"].min())


ref = os.path.join(ROOT, "oracle", "_ref")
if os.path.exists(os.path.join(ref, "pinder_0.pt")):
    ck = torch.load(os.path.join(ref, "pinder_0.pt"), weights_only=False)
    model = Score_Model(ck["state_dict"], ck["hparams"], precision="fp16").to("cuda")
    batch = batch_from_record(torch.load(os.path.join(ref, "db5_1QA9.pt"), weights_only=False), pos_width=model.pos_width)
    ms, rate, emin = timed(model, batch, 40, 40, use_clash_force=True, centre_mode=1)
    print("c2  1QA9 N=197, 40 traj x 40 steps, pinder_0, clash force: %.1f ms  %.0f pose-steps/s  (%.1f us per lock-step step)  best energy %.2f" % (ms, rate, ms * 1e3 / 41, emin))
    ms, rate, emin = timed(model, batch, 256, 40, use_clash_force=True, centre_mode=1)
    print("c2' 1QA9 N=197, 256 traj x 40 steps: %.1f ms  %.0f pose-steps/s" % (ms, rate))
sd, hp = synthetic_state_dict(0, 66), synthetic_hparams(66)
model = Score_Model(sd, hp, precision="fp16").to("cuda")
for (n, T, S, tag) in ((150, 256, 100, "c3"), (400, 64, 100, "c4"), (400, 256, 20, "c4'")):
    batch = synthetic_complex(n, n, seed=0, pos_width=66)
    ms, rate, emin = timed(model, batch, T, S)
    print("%s  synthetic 2x%d, %d traj x %d steps: %.1f ms  %.0f pose-steps/s" % (tag, n, T, S, ms, rate))
