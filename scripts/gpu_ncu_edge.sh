set -x
mkdir -p gpurun_out
PROFILE_FORWARDS=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_edge_ws -s 2 -c 1 -o gpurun_out/prof_edge_ws python profiles/run_edge_profile.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
