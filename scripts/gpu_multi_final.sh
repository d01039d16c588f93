set -x
mkdir -p gpurun_out
G=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 50 --warmup 3 > gpurun_out/bench_${G}gpu.json 2> gpurun_out/bench_${G}gpu.err; echo rc=$?
cut -c1-260 gpurun_out/bench_${G}gpu.json; tail -2 gpurun_out/bench_${G}gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $G --steps 4 --warmup 1 > gpurun_out/bench_ref_${G}gpu.json 2> gpurun_out/bench_ref_${G}gpu.err; echo rc=$?
timeout 300 python profiles/run_db5_set.py > gpurun_out/db5_c5_1gpu.log 2>&1; tail -1 gpurun_out/db5_c5_1gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 profiles/run_db5_set.py > gpurun_out/db5_c5_${G}gpu.log 2>&1; echo rc=$?; tail -1 gpurun_out/db5_c5_${G}gpu.log
python - <<PY
import csv
a = {r["id"]: r for r in csv.DictReader(open("gpurun_out/db5_c5_1gpu.csv"))}
b = {r["id"]: r for r in csv.DictReader(open("gpurun_out/db5_c5_${G}gpu.csv"))}
print("identical on 1 vs $G GPUs:", sum(a[k]["energy_checksum"] == b[k]["energy_checksum"] and a[k]["index"] == b[k]["index"] for k in a), "of", len(a))
PY
