# bench line + config #5 on $1 GPUs of one box (driver launch form), c5 result identity against the 1-GPU run
set -x
mkdir -p gpurun_out
G=${1:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/bench_${G}gpu.json 2> gpurun_out/bench_${G}gpu.err; echo rc=$?
cut -c1-200 gpurun_out/bench_${G}gpu.json; tail -2 gpurun_out/bench_${G}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${G}gpu.json"))
print("N=%d weak %.0f (%.3f ms/step) strong %s e2e %.0f" % (d["n_gpus"], d["value"], d["ms_per_step"], d.get("strong_scaling"), d["e2e"]["value"]))
PY
timeout 300 python profiles/run_db5_set.py > gpurun_out/db5_c5_1gpu.log 2>&1; tail -1 gpurun_out/db5_c5_1gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 profiles/run_db5_set.py > gpurun_out/db5_c5_${G}gpu.log 2>&1; echo rc=$?; tail -1 gpurun_out/db5_c5_${G}gpu.log
python - <<PY
import csv
a = {r["id"]: r for r in csv.DictReader(open("gpurun_out/db5_c5_1gpu.csv"))}
b = {r["id"]: r for r in csv.DictReader(open("gpurun_out/db5_c5_${G}gpu.csv"))}
print("identical on 1 vs $G GPUs:", sum(a[k]["energy_checksum"] == b[k]["energy_checksum"] and a[k]["index"] == b[k]["index"] for k in a), "of", len(a))
PY
