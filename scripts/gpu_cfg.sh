set -x
mkdir -p gpurun_out
timeout 90 python profiles/variant_check.py > gpurun_out/variant.log 2>&1; tail -1 gpurun_out/variant.log | cut -c1-150
timeout 200 python profiles/config_timings.py > gpurun_out/config_timings.log 2>&1; cat gpurun_out/config_timings.log
