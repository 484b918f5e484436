#!/bin/bash
# cluster FPS: slot pairs in registers (PN2_FPS_RP; 20 = all, two CTAs per SM; fewer = three CTAs per SM) x batches in flight
run() { echo -n "$* $LANES: "; env "$@" timeout 400 python bench.py --steps 48 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs --no-e2e $LANES 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['parity']['ok'])"; }
for rp in 4 6 8 0; do
  LANES="--lanes 12" run PN2_FPS_RP=$rp
done
LANES="--lanes 16" run PN2_FPS_RP=6
LANES="--lanes 16" run PN2_FPS_RP=8
LANES="--lanes 10" run PN2_FPS_RP=6
LANES="--lanes 12" run PN2_FPS_RP=6 PN2_SAMPLING_PRIORITY=-1
export PN2_FPS_BUCKET_MIN=1000000000
for rp in 4 6 8; do for b in 8 48; do echo -n "RP=$rp "; PN2_FPS_RP=$rp python scripts/fps_sat_one.py $b | tail -1; done; done
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "furthest or fps" 2>&1 | tail -2
