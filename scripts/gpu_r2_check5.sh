#!/bin/bash
# round-2 check 5: stream priorities and tiles-per-CTA sweep on the default bench step
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['parity']['ok'])"; }
run PN2_SAMPLING_PRIORITY=0
run PN2_SAMPLING_PRIORITY=-1
run PN2_SAMPLING_PRIORITY=-1 PN2_SA_TC_MIN_TILES=4
run PN2_SAMPLING_PRIORITY=-1 PN2_SA_TC_MIN_TILES=12
run PN2_SAMPLING_PRIORITY=0 PN2_SA_TC_MIN_TILES=6
run PN2_SAMPLING_PRIORITY=0
