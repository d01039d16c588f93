set -x
mkdir -p gpurun_out
touch dfmdock_b200/csrc/node_t.cu; DFM_NVCC_EXTRA="-DNTT_TIMING=1" python -m dfmdock_b200.build > /dev/null 2>&1
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-full-job --no-other-configs > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
grep "ntt timing" gpurun_out/bench_t.err | tail -6
