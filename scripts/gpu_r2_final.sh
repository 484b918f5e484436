#!/bin/bash
# refresh of the committed evidence at HEAD: full default bench (the line the driver prints), GPU tests, the launch
# list and the full ncu capture of one step (the other kernels' captures: scripts/gpu_profile_r2.sh; together they
# exceed the 64 MiB that travel back from the box in one call)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_full.json
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
B="python bench.py --warmup 3 --lanes 1 --no-graphs --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e --no-sub-configs --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $B --steps 2 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
PN2_FP_TC2=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"fps_kernel|sa_tc_v3|lin_tc|fp_tc_kernel|bq_grid_query|ball_query_kernel|three_nn_kernel" -s 133 -c 19 -o gpurun_out/prof_step_r2 -f $B --steps 1 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
ls -la gpurun_out
