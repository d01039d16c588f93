set -x
mkdir -p gpurun_out
touch dfmdock_b200/csrc/edge_ws.cu; DFM_NVCC_EXTRA="-DEWS_TIMING=1 ${EXTRA}" python -m dfmdock_b200.build > /dev/null 2>&1
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-full-job --no-other-configs > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
grep "ews timing" gpurun_out/bench_t.err | tail -3
