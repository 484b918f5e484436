#!/bin/bash
# Round-1 evidence run: full ncu capture of the step's kernels (last bench step, lanes=1; 19 of this library's launches per step match the filter).
mkdir -p gpurun_out
B="python bench.py --warmup 3 --lanes 1 --no-graphs --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"fps_kernel|sa_tc_v3|lin_tc|fp_tc_kernel|bq_grid_query|ball_query_kernel|three_nn_kernel" -s 133 -c 19 -o gpurun_out/prof_step_r1 -f $B --steps 1 > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/ncu_step.log
ls -la gpurun_out | grep -E "prof_step|launches_r1"
