set -x
mkdir -p gpurun_out
touch dfmdock_b200/csrc/edge_ws.cu
DFM_NVCC_EXTRA="-DEWS_TIMING=1" python -m dfmdock_b200.build > /dev/null 2>&1
timeout 90 python profiles/variant_check.py > gpurun_out/timing.log 2>&1
grep -E "ews timing|edge kernel" gpurun_out/timing.log | cut -c1-330
