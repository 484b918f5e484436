#!/bin/bash
# the driver's invocation is --steps 20 --warmup 5: pipeline fill and drain are inside the timed region, so the best
# number of batches in flight for 20 steps need not be the steady-state optimum
run() { echo -n "$ARGS: "; timeout 900 python bench.py --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs --no-e2e $ARGS 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['parity']['ok'])"; }
for l in 8 10 12 16 20; do ARGS="--steps 20 --warmup 5 --lanes $l" run; done
for l in 10 20; do ARGS="--steps 20 --warmup 5 --lanes $l" run; done
for l in 8 16; do ARGS="--steps 64 --warmup 3 --lanes $l" run; done
