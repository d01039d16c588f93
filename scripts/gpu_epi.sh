set -x
mkdir -p gpurun_out
rm -f gpurun_out/epi.log
for v in "-DEWS_EPI_PIPE=0 -DEWS_EPI_CONST=0" "-DEWS_EPI_PIPE=1 -DEWS_EPI_CONST=0" "-DEWS_EPI_PIPE=1 -DEWS_EPI_CONST=1"; do
  touch dfmdock_b200/csrc/edge_ws.cu
  DFM_NVCC_EXTRA="$v" python -m dfmdock_b200.build > /dev/null 2>&1
  timeout 120 python profiles/variant_check.py 2>&1 | grep "edge kernel" | sed "s/^/[$v] /" | cut -c1-400 >> gpurun_out/epi.log
done
cat gpurun_out/epi.log
touch dfmdock_b200/csrc/edge_ws.cu
python -m dfmdock_b200.build > /dev/null 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
