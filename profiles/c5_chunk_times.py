"""Device time of every (complex, trajectory count) work item of BASELINE config #5 on ONE GPU: the data the cost model of
dfmdock_b200.distributed.plan_work is fitted to and checked against (profiles/r02/c5_chunk_times.txt).  For each db5 complex:
set_complex + sample(T trajectories x 40 steps, clash force) + the device-to-host copy of the results, T in {40, 20, 10, 5}."""
import glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dfmdock_b200 import Score_Model
from dfmdock_b200.features import batch_from_record

ref = os.path.join(ROOT, "oracle", "_ref")
ck = torch.load(os.path.join(ref, "pinder_0.pt"), weights_only=False)
model = Score_Model(ck["state_dict"], ck["hparams"], precision="fp16").to("cuda")
paths = sorted(glob.glob(os.path.join(ref, "db5_all", "*.pt"))) or sorted(glob.glob(os.path.join(ref, "db5_*.pt")))
for p in paths:
    cid = os.path.splitext(os.path.basename(p))[0].replace("db5_", "")
    batch = batch_from_record(torch.load(p, weights_only=False), pos_width=model.pos_width, with_position_matrix=False)
    N = batch["rec_pos"].shape[0] + batch["lig_pos"].shape[0]
    out = []
    for T in (40, 20, 10, 5):
        best = 1e9
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.set_complex(batch)
            res = model.sample(batch["lig_pos"], T, num_steps=40, use_clash_force=True, centre_mode=1, seed=42)
            host = {k: v.cpu() for k, v in res.items()}
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out.append("T=%d %.2f ms" % (T, best))
    print("CHUNK %s N=%d  %s" % (cid, N, "  ".join(out)), flush=True)
