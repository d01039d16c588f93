# The CSV entry point under torchrun (planned sharding) against the same command on one GPU.
set -x
mkdir -p gpurun_out /tmp/cli
python - <<'PY'
import os, torch, sys
sys.path.insert(0, ".")
from dfmdock_b200.synthetic import write_lightning_ckpt
ck = torch.load("oracle/_ref/pinder_0.pt", weights_only=False)
write_lightning_ckpt("/tmp/cli/pinder.ckpt", ck["state_dict"], ck["hparams"])
with open("/tmp/cli/list.csv", "w") as f:
    for cid in ("1QA9", "7CEI", "4POU"):
        p = os.path.abspath("oracle/_ref/db5_%s.pt" % cid)
        f.write("%s,%s,%s\n" % (cid, p, p))
PY
ARGS="--csv /tmp/cli/list.csv --ckpt /tmp/cli/pinder.ckpt --num_samples 12 --num_steps 6 --use_clash_force --seed 7"
timeout 300 python -m dfmdock_b200.inference $ARGS --out_dir /tmp/cli/pdb1 --out_csv_dir /tmp/cli --out_csv one.csv > gpurun_out/cli_1gpu.log 2>&1; echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 -m dfmdock_b200.inference $ARGS --out_dir /tmp/cli/pdb2 --out_csv_dir /tmp/cli --out_csv two.csv > gpurun_out/cli_2gpu.log 2>&1; echo rc=$?
tail -3 gpurun_out/cli_2gpu.log
python - <<'PY'
import csv
a = list(csv.DictReader(open("/tmp/cli/one.csv"))); b = list(csv.DictReader(open("/tmp/cli/two.csv")))
print("rows", len(a), len(b), "identical:", a == b)
import os
print("pdb files", len(os.listdir("/tmp/cli/pdb1")), len(os.listdir("/tmp/cli/pdb2")))
if a != b:
    for x, y in zip(a, b):
        if x != y: print(x, y); break
PY
