#!/bin/bash
mkdir -p gpurun_out
export PN2_FPS_BUCKET_MIN=1000000000
for rp in 20 6 0; do for b in 8 32 48; do echo -n "RP=$rp "; PN2_FPS_RP=$rp timeout 120 python scripts/fps_sat_one.py $b | tail -1; done; done
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "furthest or fps" 2>&1 | tail -2
for rp in 20 6; do
  echo "== bench PN2_FPS_RP=$rp"
  PN2_FPS_RP=$rp timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['parity']['ok'])"
done
