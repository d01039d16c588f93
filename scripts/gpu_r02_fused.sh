# round 2: last layer: parity tests, then bench lines unfused / fused (edge role + coordinate-head role in one launch)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x > gpurun_out/pytest_fused.log 2>&1; tail -3 gpurun_out/pytest_fused.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-full-job --no-other-configs"
for v in "DFM_LAST_FUSED=0" "DFM_LAST_FUSED=1 DFM_LAST_FUSED_SPLIT=58" "DFM_LAST_FUSED=0"; do
  env $v timeout 300 python bench.py $B > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err || tail -5 gpurun_out/bench_f.err
  python - "$v" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_f.json"))
    print("AB [%s] value %.0f ms/step %.3f edge ms %.4f frac %.3f e2e %.0f launches %d" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"]))
except Exception as e:
    print("AB [%s] failed" % sys.argv[1], e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_coord|k_edge_ws<1>|k_edge_wsILb1" -c 6 --csv --log-file gpurun_out/launches_last.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-job --no-other-configs > gpurun_out/ncu_last.log 2>&1
grep -E "k_coord|k_edge" gpurun_out/launches_last.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
