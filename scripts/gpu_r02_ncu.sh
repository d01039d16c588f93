# round 2: re-run the tests that changed, ncu --set full of every kernel of one forward, launch list of config #2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_pdbio.py -m gpu -q -s -rA > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
grep -E "passed|failed|rc=|worst|T4 |FAILED|Error|assert" gpurun_out/pytest_gpu2.log | cut -c1-700 | tail -40
PROFILE_FORWARDS=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_nodeT|k_coord|k_graph|k_edge_ws|k_force|k_prepare|k_broadcast" -c 40 -o gpurun_out/prof_fwd python profiles/run_edge_profile.py > gpurun_out/ncu_fwd.log 2>&1
tail -2 gpurun_out/ncu_fwd.log; ls -la gpurun_out/prof_fwd.ncu-rep
PROFILE_CONFIG=c2 PROFILE_FORWARDS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python profiles/run_edge_profile.py > gpurun_out/ncu_c2.log 2>&1
tail -2 gpurun_out/ncu_c2.log
