set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "forward_injected or without_energy or batched_equals or real_checkpoints" > gpurun_out/pytest_quick.log 2>&1; tail -2 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_a.json"))
print("value %.0f ms/step %.3f edge ms %.4f frac %.3f e2e %.0f full_job %.0f c2 us/step %.0f c4 %.0f launches %d" % (
    d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["e2e"]["value"],
    d["full_job"]["poses_per_s"], d["other_configs"]["c2"]["us_per_lockstep_step"], d["other_configs"]["c4"]["poses_per_s"], d["gpu_launches"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-job --no-other-configs > gpurun_out/ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_bench.csv > gpurun_out/launches_bench_summary.txt 2>&1; cat gpurun_out/launches_bench_summary.txt
