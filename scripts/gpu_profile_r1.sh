#!/bin/bash
# Round-1 evidence run: GPU parity suite, smoke, ncu launch list + full captures of the top kernels, bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
B="python bench.py --warmup 3 --lanes 1 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1.csv $B --steps 2 > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_kernel -s 12 -c 4 -o gpurun_out/prof_fps_r1 -f $B --steps 1 > gpurun_out/ncu_fps.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_tc -s 12 -c 4 -o gpurun_out/prof_satc_r1 -f $B --steps 1 > gpurun_out/ncu_satc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fp_tc_kernel|bq_grid_query" -s 6 -c 3 -o gpurun_out/prof_fp_bq_r1 -f $B --steps 1 > gpurun_out/ncu_fpbq.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_r1.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_r1.log
timeout 600 python bench.py --steps 30 --warmup 3 --lanes 1 --no-cpu-baseline --no-reference-cuda --no-kernel-breakdown > gpurun_out/bench_r1_lanes1.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline --no-reference-cuda > gpurun_out/bench_r1_fp32.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 1 --impl reference > gpurun_out/bench_r1_ref.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
ls -la gpurun_out | head -40
