#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_fused_gpu.py -m gpu -q --timeout 300 -x -k "fps or furthest or sampl or backbone or reference" > gpurun_out/pytest_fps.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fps.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/pytest_fps.log | head
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/pytest_gpu.log | head
bash scripts/gpu_lanes.sh
