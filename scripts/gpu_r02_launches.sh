set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-job --no-other-configs > gpurun_out/ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_bench.csv > gpurun_out/launches_bench_summary.txt 2>&1; cat gpurun_out/launches_bench_summary.txt
