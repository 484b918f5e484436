#!/bin/bash
# the driver's multi-GPU invocation at N = 8, both arms
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "rc=$?"
tail -1 gpurun_out/bench_8gpu.json | cut -c1-400; tail -3 gpurun_out/bench_8gpu.err | cut -c1-300
