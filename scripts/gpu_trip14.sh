#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/sa_profile.py > gpurun_out/sa_profile.log 2>&1; cat gpurun_out/sa_profile.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -20
grep -B5 -A25 "Error\|assert " gpurun_out/pytest_gpu.log | head -60
