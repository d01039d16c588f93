set -x
mkdir -p gpurun_out
timeout 90 python profiles/variant_check.py > gpurun_out/variant.log 2>&1; echo rc=$?; tail -2 gpurun_out/variant.log
grep -q "edge kernel" gpurun_out/variant.log || exit 1
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python profiles/run_edge_profile.py > gpurun_out/ncu_launches.log 2>&1
