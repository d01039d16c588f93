# round 2: compute-sanitizer memcheck + initcheck over the paths that changed this round (TMA tensor-map / gather4 loads, dependent
# launches, k_graph_big, K-block-major spill incl. the odd-batch tail, fused last-layer launch) at small shapes
set -x
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as g
g.smoke()
from dfmdock_b200 import Score_Model
from dfmdock_b200.features import synthetic_complex
from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
m = Score_Model(synthetic_state_dict(1, 66), synthetic_hparams(66)).to("cuda")
# odd N / odd number of ligand residues in the batch (spill tail), ragged graph (N < 60), clash force
for (r, l, T) in ((25, 20, 3), (41, 31, 3), (150, 151, 3)):
    bb = synthetic_complex(r, l, seed=2)
    m.set_complex(bb)
    res = m.sample(bb["lig_pos"], T, num_steps=3, seed=1, use_clash_force=True, centre_mode=1)
    print(r, l, float(res["energy"].sum()))
# tiny / ragged complexes: N < 20 (no sampled edges), one-residue chains
for (r, l) in ((8, 4), (18, 1), (1, 25), (2, 2)):
    bb = synthetic_complex(r, l, seed=21)
    m.set_complex(bb)
    o = m.score(bb["lig_pos"][None].repeat(3, 1, 1, 1).contiguous(), torch.tensor([0.8, 0.5, 0.2]), seed=7, forward_index=1, want_energy=True)
    res = m.sample(bb["lig_pos"], 2, num_steps=2, seed=1, use_clash_force=True)
    print("tiny", r, l, float(o["energy"].sum()), float(res["energy"].sum()))
# fused last layer (needs >= 4 x SMs ligand tiles): 40 trajectories of 2 x 40 residues
bb = synthetic_complex(40, 40, seed=3)
bb["lig_pos"] = bb["lig_pos"] - torch.tensor([12.0, 0.0, 0.0])
m.set_complex(bb)
lig = bb["lig_pos"][None].repeat(40, 1, 1, 1).contiguous()
t = torch.full((40,), 0.4)
a = m.score(lig, t, seed=1, forward_index=0)
m.last_fused = True
n0 = m.launch_count
b = m.score(lig, t, seed=1, forward_index=0)
print("fused launches", m.launch_count - n0, "equal", bool(torch.equal(a["f"], b["f"])))
m.last_fused = False
# large complex: k_graph_big
bb = synthetic_complex(700, 400, seed=4)
m.set_complex(bb)
o = m.score(bb["lig_pos"][None], torch.full((1,), 0.5), seed=2, forward_index=0, want_energy=True)
print("N=1100 energy", float(o["energy"][0]))
torch.cuda.synchronize()
print("sanitize script done")
PY
for tool in memcheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --launch-timeout 0 --print-limit 10 python /tmp/san2.py > gpurun_out/sanitize_r02_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|Uninitialized|done|smoke ok|fused launches|N=1100|tiny" gpurun_out/sanitize_r02_$tool.log | sort | uniq -c | sort -rn | head -12
done
