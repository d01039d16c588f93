set -x
mkdir -p gpurun_out
timeout 300 python tests/gpu_report.py > gpurun_out/gpu_report.log 2>&1; tail -40 gpurun_out/gpu_report.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python profiles/variant_check.py > gpurun_out/variant.log 2>&1; tail -3 gpurun_out/variant.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
