#!/bin/bash
# round-2 check 3: FPS register/shared split sweep, split ball-query grid test, short bench at each FPS variant
mkdir -p gpurun_out
timeout 600 python scripts/fps_rp_sweep.py 2>&1 | tee gpurun_out/fps_rp_sweep.jsonl
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "ball_query or furthest or fps" 2>&1 | tail -3
for rp in 20 8 0; do
  echo "== bench PN2_FPS_RP=$rp"
  PN2_FPS_RP=$rp timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs 2>&1 | tail -1 | tee gpurun_out/bench_check3_rp$rp.json | cut -c1-400
done
