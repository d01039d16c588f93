# round 2: full GPU test suite (worst errors printed), the bench line, the launch list of the bench command
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -rA > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|rc=|worst|T4 |KS |FAILED|Error|assert" gpurun_out/pytest_gpu.log | cut -c1-400 | tail -60
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-job --no-other-configs > gpurun_out/ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_bench.csv > gpurun_out/launches_bench_summary.txt 2>&1; cat gpurun_out/launches_bench_summary.txt
