set -x
mkdir -p gpurun_out
cat > /tmp/gb.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from dfmdock_b200 import Score_Model
from dfmdock_b200.features import synthetic_complex
from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
n, T = 1000, 16
sd, hp = synthetic_state_dict(0, 66), synthetic_hparams(66)
model = Score_Model(sd, hp, precision="fp16").to("cuda")
batch = synthetic_complex(n, n, seed=0, pos_width=model.pos_width)
model.set_complex(batch)
lig = batch["lig_pos"][None].repeat(T, 1, 1, 1).cuda().contiguous()
t = torch.full((T,), 0.3, device="cuda")
for i in range(2): model.score(lig, t, seed=0, forward_index=i)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_graph_big -s 1 -c 1 -o gpurun_out/prof_graph_big python /tmp/gb.py > gpurun_out/ncu_gb.log 2>&1; tail -2 gpurun_out/ncu_gb.log
