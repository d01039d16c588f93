set -x
mkdir -p gpurun_out
timeout 90 python profiles/variant_check.py > gpurun_out/variant.log 2>&1; echo rc=$?; tail -1 gpurun_out/variant.log
grep -q "edge kernel" gpurun_out/variant.log || exit 1
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
PROFILE_FORWARDS=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_edge_ws -s 2 -c 1 -o gpurun_out/prof_edge_ws python profiles/run_edge_profile.py > gpurun_out/ncu_full.log 2>&1
