# usage: bash scripts/gpu_var.sh "<nvcc defines variant 1>" "<variant 2>" ...   (each built on the box, timed + accuracy-checked)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/var.log
for v in "$@"; do
  touch dfmdock_b200/csrc/edge_ws.cu
  DFM_NVCC_EXTRA="$v" python -m dfmdock_b200.build > /dev/null 2>&1
  timeout 120 python profiles/variant_check.py 2>&1 | grep "edge kernel" | sed "s/^/[$v] /" | cut -c1-420 >> gpurun_out/var.log
done
cat gpurun_out/var.log
