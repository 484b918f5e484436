#!/bin/bash
# throughput-mode step: SA tiles per CTA and stream priorities once more (the balance of the step changed)
run() { echo -n "$* $ARGS: "; env "$@" timeout 600 python bench.py --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs --no-e2e $ARGS 2>gpurun_out/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['parity']['ok'])" || tail -3 gpurun_out/err.txt; }
mkdir -p gpurun_out
for v in 8 4 16 8; do ARGS="--steps 20 --warmup 5" run PN2_SA_TC_MIN_TILES=$v; done
ARGS="--steps 20 --warmup 5" run PN2_SAMPLING_PRIORITY=-1
ARGS="--steps 20 --warmup 5" run PN2_LANE_PRIORITY=-1
ARGS="--steps 20 --warmup 5" run PN2_FP_TC2=1
ARGS="--steps 20 --warmup 5" run PN2_SA_TC_TMA=0
