set -x
mkdir -p gpurun_out
for v in "default" "-DEWS_EXP=16"; do
  if [ "$v" != "default" ]; then touch dfmdock_b200/csrc/edge_ws.cu; DFM_NVCC_EXTRA="$v" python -m dfmdock_b200.build > /dev/null 2>&1; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_edge_ws" -c 12 --csv --log-file gpurun_out/launches_last.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-full-job --no-other-configs > gpurun_out/ncu_last.log 2>&1
  echo "variant $v"; grep -E "k_edge" gpurun_out/launches_last.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
done
