set -x
mkdir -p gpurun_out
PROFILE_FORWARDS=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_tc<\(int\)1' -s 2 -c 1 -o gpurun_out/prof_edge python profiles/run_edge_profile.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
DFM_EDGE_VARIANT=0 timeout 300 python profiles/variant_check.py > gpurun_out/variant0.log 2>&1
DFM_EDGE_VARIANT=1 timeout 300 python profiles/variant_check.py > gpurun_out/variant1.log 2>&1
tail -2 gpurun_out/variant0.log gpurun_out/variant1.log
