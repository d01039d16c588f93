set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python tests/gpu_report.py > gpurun_out/gpu_report.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python profiles/run_edge_profile.py > gpurun_out/ncu_launches.log 2>&1
PROFILE_FORWARDS=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_tc<1" -s 2 -c 2 -o gpurun_out/prof_edge python profiles/run_edge_profile.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
