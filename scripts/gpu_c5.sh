set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 400 python profiles/run_db5_set.py > gpurun_out/db5_c5_1gpu.log 2>&1; echo rc=$?; tail -3 gpurun_out/db5_c5_1gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 profiles/run_db5_set.py > gpurun_out/db5_c5_2gpu.log 2>&1; echo rc=$?; tail -3 gpurun_out/db5_c5_2gpu.log
python - <<'PY'
import csv
a = {r["id"]: r for r in csv.DictReader(open("gpurun_out/db5_c5_1gpu.csv"))}
b = {r["id"]: r for r in csv.DictReader(open("gpurun_out/db5_c5_2gpu.csv"))}
same = sum(a[k]["energy_checksum"] == b[k]["energy_checksum"] and a[k]["index"] == b[k]["index"] for k in a)
print("sharding invariance: %d of %d complexes have identical energy checksums and best sample on 1 and 2 GPUs" % (same, len(a)))
for k in a:
    if a[k]["energy_checksum"] != b[k]["energy_checksum"]:
        print("  differs:", k, a[k]["energy_checksum"], b[k]["energy_checksum"], a[k]["chunks"], "|", b[k]["chunks"])
PY
