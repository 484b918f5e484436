#!/bin/bash
# 2-GPU runs: bench (scene-sharded, no collective), training step (config 4: NCCL gradient all-reduce), gloo tests' NCCL twin
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-kernel-breakdown > gpurun_out/bench_2gpu.log 2>&1
grep '^{' gpurun_out/bench_2gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('2 GPUs: value %.0f scenes/s, e2e %.0f' % (d['value'], d['e2e']['value']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 scripts/train_bench.py > gpurun_out/train_2gpu.log 2>&1; grep '^{' gpurun_out/train_2gpu.log || tail -5 gpurun_out/train_2gpu.log
timeout 600 python scripts/train_bench.py > gpurun_out/train_1gpu.log 2>&1; grep '^{' gpurun_out/train_1gpu.log || tail -5 gpurun_out/train_1gpu.log
