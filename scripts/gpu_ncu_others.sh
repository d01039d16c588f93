set -x
mkdir -p gpurun_out
PROFILE_FORWARDS=1 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_nodeT|k_node<|k_graph_sel" -s 0 -c 24 -o gpurun_out/prof_others python profiles/run_edge_profile.py > gpurun_out/ncu_others.log 2>&1
tail -2 gpurun_out/ncu_others.log
