set -x
mkdir -p gpurun_out
PROFILE_FORWARDS=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_node -s 10 -c 4 -o gpurun_out/prof_node python profiles/run_edge_profile.py > gpurun_out/ncu_node.log 2>&1
tail -3 gpurun_out/ncu_node.log
