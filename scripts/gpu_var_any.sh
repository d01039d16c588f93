# usage: bash scripts/gpu_var_any.sh <file.cu to touch> "<defines 1>" "<defines 2>" ...
set -x
mkdir -p gpurun_out
rm -f gpurun_out/var.log
f=$1; shift
for v in "$@"; do
  touch dfmdock_b200/csrc/$f
  DFM_NVCC_EXTRA="$v" python -m dfmdock_b200.build > /dev/null 2>&1
  timeout 120 python profiles/variant_check.py 2>&1 | grep "edge kernel" | sed "s/^/[$v] /" | cut -c1-420 >> gpurun_out/var.log
done
cat gpurun_out/var.log
