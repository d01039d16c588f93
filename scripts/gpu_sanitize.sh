# compute-sanitizer over the smoke path (small complex: forward with energy, reverse step, 4 x 3 sampling) and the metrics / atoms kernels
set -x
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as g
g.smoke()
from dfmdock_b200 import pdbio
from dfmdock_b200.metrics import compute_metrics_batch
from dfmdock_b200.features import synthetic_complex
b = synthetic_complex(33, 27, seed=1)
lig = b["lig_pos"][None].repeat(3, 1, 1, 1) + 1.0
print(compute_metrics_batch(b["rec_pos"], lig, b["rec_pos"], b["lig_pos"], device="cuda")[0])
print(pdbio.modify_aa_coords(torch.randn(50, 3), b["lig_pos"], torch.randn(3, 3), torch.randn(3, 3), device="cuda").shape)
# odd N, ragged graph (N < 60), ligand-only last layer, clash force
from dfmdock_b200 import Score_Model
from dfmdock_b200.synthetic import synthetic_hparams, synthetic_state_dict
m = Score_Model(synthetic_state_dict(1, 66), synthetic_hparams(66)).to("cuda")
for (r, l) in ((25, 20), (41, 30), (150, 151)):
    bb = synthetic_complex(r, l, seed=2)
    m.set_complex(bb)
    res = m.sample(bb["lig_pos"], 3, num_steps=3, seed=1, use_clash_force=True, centre_mode=1)
    print(r, l, float(res["energy"].sum()))
torch.cuda.synchronize()
print("sanitize script done")
PY
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 0 --print-limit 20 python /tmp/san.py > gpurun_out/sanitize_memcheck.log 2>&1; echo rc=$?
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|done|smoke ok" gpurun_out/sanitize_memcheck.log | head -20
