# A/B of node-kernel build switches: quick parity tests, then per-kernel times from the ncu launch list of the bench command
set -x
mkdir -p gpurun_out
IFS=';' read -ra VS <<< "${VARIANTS:-default}"
for v in "${VS[@]}"; do
  if [ "$v" != "default" ]; then touch dfmdock_b200/csrc/*.cu; DFM_NVCC_EXTRA="$v" python -m dfmdock_b200.build > /dev/null 2>&1; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "forward_injected or without_energy or batched_equals or real_checkpoints" > gpurun_out/pytest_quick.log 2>&1; tail -1 gpurun_out/pytest_quick.log
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-job --no-other-configs > gpurun_out/ncu_launches.log 2>&1
  echo "VARIANT [$v]"; python scripts/launch_summary.py gpurun_out/launches_bench.csv | grep -E "nodeT|graphnorm|total"
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-full-job --no-other-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   bench value %.0f ms/step %.3f' % (d['value'], d['ms_per_step']))"
done
