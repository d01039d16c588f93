set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/pytest_fused.log 2>&1; tail -2 gpurun_out/pytest_fused.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_coord|k_edge_ws" -c 12 --csv --log-file gpurun_out/launches_last.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-job --no-other-configs > gpurun_out/ncu_last.log 2>&1
grep -E "k_coord|k_edge_ws<1>" gpurun_out/launches_last.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
