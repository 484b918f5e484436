#!/bin/bash
# round-2 check 4: restructured grid ball query (tests + kernel time), SA grid-size heuristic A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_fused_gpu.py -x -q -m gpu 2>&1 | tail -3
B="python bench.py --warmup 3 --lanes 1 --no-graphs --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e --no-sub-configs --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_check4.csv $B --steps 2 > gpurun_out/ncu_launches4.log 2>&1
grep -E "bq_grid_query|sa_tc_v3" gpurun_out/launches_check4.csv | tail -5 | awk -F'","' '{print $5, $9, $NF}' | cut -c1-200
for mt in 1 4 8; do
  echo "== bench PN2_SA_TC_MIN_TILES=$mt"
  PN2_SA_TC_MIN_TILES=$mt timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_check4_mt$mt.json | cut -c1-330
done
