set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "distribution" > gpurun_out/pytest_stat.log 2>&1; tail -8 gpurun_out/pytest_stat.log
