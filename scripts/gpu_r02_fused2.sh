set -x
mkdir -p gpurun_out
for sp in 45 52 58 68; do
DFM_LAST_FUSED_SPLIT=$sp PROFILE_FORWARDS=2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_last_fused -c 2 --csv --log-file gpurun_out/launches_fused_$sp.csv python profiles/run_edge_profile.py > gpurun_out/ncu_fused.log 2>&1
echo "split $sp"; grep "k_last_fused" gpurun_out/launches_fused_$sp.csv | awk -F'","' '{print $13, $15}' | tr -d '"'
done
