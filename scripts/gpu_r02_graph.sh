set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "graph or large_complex or knn or generic" > gpurun_out/pytest_graph.log 2>&1; tail -5 gpurun_out/pytest_graph.log
python - <<'PY'
import os, subprocess, sys
code = '''
import sys, torch
sys.path.insert(0, ".")
from dfmdock_b200 import Score_Model
from dfmdock_b200.features import synthetic_complex
from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
from torch.profiler import profile, ProfilerActivity
n, T = int(sys.argv[1]), int(sys.argv[2])
sd, hp = synthetic_state_dict(0, 66), synthetic_hparams(66)
model = Score_Model(sd, hp, precision="fp16").to("cuda")
batch = synthetic_complex(n, n, seed=0, pos_width=model.pos_width)
model.set_complex(batch)
lig = batch["lig_pos"][None].repeat(T, 1, 1, 1).cuda().contiguous()
t = torch.full((T,), 0.3, device="cuda")
for i in range(3): model.score(lig, t, seed=0, forward_index=i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10): model.score(lig, t, seed=0, forward_index=3 + i)
e1.record(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(5): model.score(lig, t, seed=0, forward_index=20 + i)
    torch.cuda.synchronize()
g = [ev for ev in prof.key_averages() if "k_graph" in ev.key]
print("RESULT %.3f %.3f %s" % (e0.elapsed_time(e1) / 10, sum(ev.device_time_total for ev in g) / 5e3, g[0].key[:24] if g else ""))
'''
for n, T in ((600, 32), (1000, 16), (1274, 10), (2000, 8)):
    for tag, env in (("big", {}), ("generic", {"DFM_GRAPH_BIG": "0"})):
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, "-c", code, str(n), str(T)], env=e, capture_output=True, text=True)
        r = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        print("N=2x%d T=%d %-8s" % (n, T, tag), r[0] if r else out.stderr[-400:])
PY
timeout 600 python profiles/run_db5_set.py > gpurun_out/db5_c5_1gpu.log 2>&1; tail -3 gpurun_out/db5_c5_1gpu.log
DFM_GRAPH_BIG=0 timeout 600 python profiles/run_db5_set.py > gpurun_out/db5_c5_1gpu_generic.log 2>&1; tail -1 gpurun_out/db5_c5_1gpu_generic.log
