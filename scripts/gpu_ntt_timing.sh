set -x
mkdir -p gpurun_out
for v in "-DNTT_TIMING=1 -DNTT_BULK_W=0" "-DNTT_TIMING=1 -DNTT_BULK_W=1"; do
touch dfmdock_b200/csrc/node_t.cu
DFM_NVCC_EXTRA="$v" python -m dfmdock_b200.build > /dev/null 2>&1
timeout 120 python profiles/variant_check.py > gpurun_out/ntt_timing.log 2>&1
echo "== $v"; grep -E "ntt timing|edge kernel" gpurun_out/ntt_timing.log | cut -c1-330 | tail -4
done
touch dfmdock_b200/csrc/node_t.cu; python -m dfmdock_b200.build > /dev/null 2>&1
timeout 120 python profiles/variant_check.py 2>&1 | tail -1 | cut -c1-200
