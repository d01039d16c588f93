set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_metrics.py -m gpu -x -q > gpurun_out/pytest_metrics.log 2>&1; tail -15 gpurun_out/pytest_metrics.log
