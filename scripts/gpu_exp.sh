set -x
mkdir -p gpurun_out
for e in 0 1 2 3 4 7; do
  touch dfmdock_b200/csrc/edge_ws.cu
  DFM_NVCC_EXTRA="-DEWS_EXP=$e" python -m dfmdock_b200.build > /dev/null 2>&1
  timeout 90 python profiles/variant_check.py 2>&1 | grep "edge kernel" | sed "s/^/EXP=$e /" | cut -c1-120 >> gpurun_out/exp.log
done
cat gpurun_out/exp.log
