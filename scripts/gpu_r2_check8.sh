#!/bin/bash
run() { echo -n "$ARGS: "; timeout 900 python bench.py --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs --no-e2e $ARGS 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['parity']['ok'])"; }
ARGS="--config 5 --lanes 24 --steps 48 --warmup 3" run
ARGS="--config 5 --lanes 32 --steps 64 --warmup 3" run
ARGS="--config 5 --points 100000 --lanes 24 --steps 48 --warmup 3" run
nvidia-smi --query-gpu=memory.total --format=csv
