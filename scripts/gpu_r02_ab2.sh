# round 2: quick tests, then A/B bench lines of edge-kernel build switches: $VARIANTS = ';'-separated nvcc -D sets (first = default build)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "forward_injected or without_energy or batched_equals or real_checkpoints" > gpurun_out/pytest_quick.log 2>&1; tail -2 gpurun_out/pytest_quick.log
B="--steps 20 --warmup 5 --no-cpu-baseline --no-full-job --no-other-configs"
IFS=';' read -ra VS <<< "${VARIANTS:-default}"
i=0
for v in "${VS[@]}"; do
  if [ "$v" != "default" ]; then touch dfmdock_b200/csrc/*.cu; DFM_NVCC_EXTRA="$v" python -m dfmdock_b200.build > /dev/null 2>&1; fi
  timeout 600 python bench.py $B > gpurun_out/bench_ab_$i.json 2> gpurun_out/bench_ab_$i.err
  python - "$v" gpurun_out/bench_ab_$i.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("AB [%s] value %.0f ms/step %.3f edge ms %.4f frac %.3f e2e %.0f" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["e2e"]["value"]))
except Exception as e:
    print("AB [%s] failed" % sys.argv[1], e)
PY
  i=$((i+1))
done
