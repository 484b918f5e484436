#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scripts/ubench > gpurun_out/ubench.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/ubench.log; tail -30 gpurun_out/pytest_gpu.log
