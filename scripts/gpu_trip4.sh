#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/fps_sweep.py profile > gpurun_out/fps_profile.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_kernel -s 4 -c 1 -o gpurun_out/prof_fps_r1 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e > gpurun_out/ncu_fps.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_tc_kernel -s 4 -c 2 -o gpurun_out/prof_satc_r1 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e > gpurun_out/ncu_satc.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_bf16.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -40
cat gpurun_out/fps_profile.log; tail -3 gpurun_out/ncu_launches.log | cut -c 1-300; tail -2 gpurun_out/ncu_fps.log | cut -c 1-300; tail -2 gpurun_out/ncu_satc.log | cut -c 1-300
ls -la gpurun_out
