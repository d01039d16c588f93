set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo rc=$?
cut -c1-260 gpurun_out/bench_8gpu.json; tail -3 gpurun_out/bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 profiles/run_db5_set.py > gpurun_out/db5_c5_8gpu.log 2>&1; echo rc=$?; tail -4 gpurun_out/db5_c5_8gpu.log
