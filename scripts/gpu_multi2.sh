set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo rc=$?
cat gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err; echo rc=$?
cut -c1-250 gpurun_out/bench_ref_2gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 profiles/run_db5_set.py > gpurun_out/db5_c5_2gpu.log 2>&1; echo rc=$?; tail -8 gpurun_out/db5_c5_2gpu.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
