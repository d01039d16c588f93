# round 2 evidence: ncu --set full of every kernel of a forward (two captures), whole-step DRAM bytes, config #4 pair-tile sweep,
# config #2 timeline, other-config timings
set -x
mkdir -p gpurun_out
PROFILE_FORWARDS=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_prepare|k_graph_sel|k_broadcast_h0|k_nodeT|k_edge_ws|k_graphnorm_stats" -c 8 -o gpurun_out/prof_fwd_a python profiles/run_edge_profile.py > gpurun_out/ncu_fwd_a.log 2>&1; tail -1 gpurun_out/ncu_fwd_a.log
PROFILE_FORWARDS=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_coord|k_force_head|k_reverse_step|k_edge_wsILb1" --kernel-name-base mangled -c 4 -o gpurun_out/prof_fwd_b python profiles/run_edge_profile.py > gpurun_out/ncu_fwd_b.log 2>&1; tail -1 gpurun_out/ncu_fwd_b.log
PROFILE_FORWARDS=2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 120 --csv --log-file gpurun_out/step_dram.csv python profiles/run_edge_profile.py > gpurun_out/ncu_dram.log 2>&1; tail -1 gpurun_out/ncu_dram.log
timeout 600 python profiles/c4_sweep.py > gpurun_out/c4_pair_tile_sweep.txt 2>&1; cat gpurun_out/c4_pair_tile_sweep.txt
timeout 300 python profiles/c2_timeline.py > gpurun_out/c2_timeline.txt 2>&1; tail -3 gpurun_out/c2_timeline.txt
timeout 300 python profiles/config_timings.py > gpurun_out/config_timings.log 2>&1; cat gpurun_out/config_timings.log
ls -la gpurun_out/*.ncu-rep
