#!/bin/bash
# batches in flight for configs 3 / 5 with the three-CTA-per-SM sampling kernel
run() { echo -n "$ARGS: "; timeout 900 python bench.py --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs --no-e2e $ARGS 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['parity']['ok'])"; }
ARGS="--config 3 --lanes 32 --steps 96" run
ARGS="--config 3 --lanes 48 --steps 96" run
ARGS="--config 5 --lanes 8 --steps 24" run
ARGS="--config 5 --lanes 12 --steps 24" run
ARGS="--config 5 --lanes 16 --steps 32" run
ARGS="--config 5 --points 100000 --lanes 12 --steps 24" run
ARGS="--config 1 --steps 20" run
nvidia-smi --query-gpu=memory.used --format=csv
