#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1
echo "exit $?" >> gpurun_out/bench_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 1 --impl reference > gpurun_out/bench_2gpu_ref.log 2>&1
echo "exit $?" >> gpurun_out/bench_2gpu_ref.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/bench_1gpu.log 2>&1
tail -c 1500 gpurun_out/bench_2gpu.log; echo; tail -c 800 gpurun_out/bench_2gpu_ref.log; echo; tail -c 600 gpurun_out/bench_1gpu.log | head -c 400
