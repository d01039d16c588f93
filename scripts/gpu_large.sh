set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_1N2C.csv python profiles/large_complex_profile.py > gpurun_out/large.log 2>&1; tail -2 gpurun_out/large.log
python scripts/launch_summary.py gpurun_out/launches_1N2C.csv | head -30
