#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -q --timeout 120 -x -k "selftest" > gpurun_out/pytest_tc_selftest.log 2>&1
echo "selftest exit $?" >> gpurun_out/pytest_tc_selftest.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/fps_sweep.py > gpurun_out/fps_sweep.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_bf16.log
tail -5 gpurun_out/pytest_tc_selftest.log; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -40
cat gpurun_out/fps_sweep.log; tail -3 gpurun_out/bench_bf16.log | cut -c 1-1500
