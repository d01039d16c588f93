# round 2: tests, then A/B bench lines: default | DFM_PDL=0 | -DEWS_SKIP_PAD=0
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -5 gpurun_out/pytest_gpu3.log
B="--steps 20 --warmup 5 --no-cpu-baseline"
timeout 600 python bench.py $B > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err
DFM_PDL=0 timeout 600 python bench.py $B > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err
touch dfmdock_b200/csrc/edge_ws.cu; DFM_NVCC_EXTRA="-DEWS_SKIP_PAD=0" python -m dfmdock_b200.build > /dev/null 2>&1
timeout 600 python bench.py $B > gpurun_out/bench_nopad.json 2> gpurun_out/bench_nopad.err
touch dfmdock_b200/csrc/edge_ws.cu; python -m dfmdock_b200.build > /dev/null 2>&1
python - <<'PY'
import json
for n in ("a", "nopdl", "nopad"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % n))
        print(n, "value %.0f ms/step %.3f edge ms %.4f frac %.3f e2e %.0f full_job %.0f c2 us/step %.0f c4 %.0f launches %d" % (
            d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["e2e"]["value"],
            d["full_job"]["poses_per_s"], d["other_configs"]["c2"]["us_per_lockstep_step"], d["other_configs"]["c4"]["poses_per_s"], d["gpu_launches"]))
    except Exception as e:
        print(n, "failed", e)
PY
