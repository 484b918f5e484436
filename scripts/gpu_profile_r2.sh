#!/bin/bash
# Round-2 evidence run.  (1) launch list of a bench step, (2) full ncu capture of the step's kernels (last bench step,
# lanes=1, eager launches; 19 of this library's launches per step match the filter), (3) the kernels added or rewritten
# this round at their bench shapes, (4) SASS opcode histogram of the library.
mkdir -p gpurun_out
B="python bench.py --warmup 3 --lanes 1 --no-graphs --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e --no-sub-configs --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $B --steps 2 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
PN2_FP_TC2=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"fps_kernel|sa_tc_v3|lin_tc|fp_tc_kernel|bq_grid_query|ball_query_kernel|three_nn_kernel" -s 133 -c 19 -o gpurun_out/prof_step_r2 -f $B --steps 1 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fp_tc2_kernel" -s 2 -c 2 -o gpurun_out/prof_fp2_r2 -f python scripts/fp_iso.py > gpurun_out/ncu_fp2.log 2>&1
tail -1 gpurun_out/ncu_fp2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fps_bucket_kernel" -s 1 -c 1 -o gpurun_out/prof_fpsb_r2 -f python scripts/fps_bucket_one.py 200000 4096 8 > gpurun_out/ncu_fpsb.log 2>&1
tail -1 gpurun_out/ncu_fpsb.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"reencode_v2_kernel|prior_kernel|linear_gelu|column_pool|token_gather" -c 8 -o gpurun_out/prof_config3_r2 -f python scripts/configs_bench.py > gpurun_out/ncu_config3.log 2>&1
tail -1 gpurun_out/ncu_config3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"projection_flags|projection_compact|project_map|project_dense|project_maxpool" -s 8 -c 6 -o gpurun_out/prof_projection_r2 -f python scripts/projection_bench.py > gpurun_out/ncu_projection.log 2>&1
tail -1 gpurun_out/ncu_projection.log
ls -la gpurun_out | grep -E "prof_.*_r2|launches_r2"
