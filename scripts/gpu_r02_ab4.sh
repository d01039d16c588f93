# round 2: A/B bench lines of prebuilt edge-kernel variants (scripts/build_variants.sh): every dfmdock_b200/lib_variants/*.so is
# swapped in for the default library in turn; $QUICK_TESTS=1 also runs the quick parity tests on each
set -x
mkdir -p gpurun_out
cp dfmdock_b200/lib/libdfmdock_b200.so /tmp/default.so
B="--steps 20 --warmup 5 --no-cpu-baseline --no-full-job --no-other-configs"
for so in /tmp/default.so $(ls dfmdock_b200/lib_variants/*.so); do
  name=$(basename $so .so)
  cp $so dfmdock_b200/lib/libdfmdock_b200.so
  if [ "${QUICK_TESTS:-0}" = "1" ]; then
    timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q -x -s -k "forward_injected or without_energy or batched_equals or real_checkpoints or oracle" > gpurun_out/pytest_$name.log 2>&1
    grep -E "worst|passed|failed" gpurun_out/pytest_$name.log | cut -c1-500 | tail -12
  fi
  timeout 600 python bench.py $B > gpurun_out/bench_ab_$name.json 2> gpurun_out/bench_ab_$name.err
  python - "$name" gpurun_out/bench_ab_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("AB [%s] value %.0f ms/step %.3f edge ms %.4f frac %.3f e2e %.0f" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["e2e"]["value"]))
except Exception as e:
    print("AB [%s] failed" % sys.argv[1], e)
PY
done
cp /tmp/default.so dfmdock_b200/lib/libdfmdock_b200.so
