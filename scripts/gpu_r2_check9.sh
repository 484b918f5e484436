#!/bin/bash
# the driver's invocation (--steps 20 --warmup 5): batches in flight x bucketed kernel as one launch / split, repeated
run() { echo -n "$* $ARGS: "; env "$@" timeout 900 python bench.py --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs --no-e2e $ARGS 2>gpurun_out/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['parity']['ok'], d['config'].get('sampling_mode'))" || tail -3 gpurun_out/err.txt; }
mkdir -p gpurun_out
for rep in 1 2 3; do
for l in 20 32; do for sp in 0 1; do
ARGS="--steps 20 --warmup 5 --lanes $l" run PN2_FPS_BUCKET_SPLIT=$sp
done; done; done
ARGS="--steps 20 --warmup 5 --lanes 24" run PN2_FPS_BUCKET_SPLIT=0
ARGS="--steps 20 --warmup 5 --lanes 24" run PN2_FPS_BUCKET_SPLIT=1
ARGS="--steps 20 --warmup 5 --lanes 20 --sampling-mode latency" run A=1
