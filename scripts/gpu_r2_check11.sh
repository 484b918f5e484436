#!/bin/bash
# small-scene FPS levels (2048 / 1024 / 512 points, one CTA per scene): threads per CTA
run() { echo -n "$* $ARGS: "; env "$@" timeout 600 python bench.py --no-cpu-baseline --no-reference-cuda --no-sub-configs --no-e2e $ARGS 2>gpurun_out/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), d['parity']['ok'], {k['kernel']: round(k['us'],1) for k in d['roofline_kernels'] if k['kernel'].startswith('fps')})" || tail -3 gpurun_out/err.txt; }
mkdir -p gpurun_out
ARGS="--config 1 --steps 20" run A=1
ARGS="--config 1 --steps 20" run PN2_FPS_THREADS=128
ARGS="--config 1 --steps 20" run PN2_FPS_THREADS=512
ARGS="--steps 20 --warmup 5" run PN2_FPS_THREADS=128
ARGS="--steps 20 --warmup 5" run A=1
