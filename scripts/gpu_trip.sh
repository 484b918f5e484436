#!/bin/bash
# One GPU-box visit: reference-kernel goldens, GPU parity suite, smoke, bench, FPS sweep.
# Every stage has its own timeout so that a hung kernel cannot hold the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tests/golden/make_ref_cuda_goldens.py > gpurun_out/goldens.log 2>&1
echo "goldens exit $?" >> gpurun_out/goldens.log
if [ ! -f tests/golden/ref_cuda_ops.npz ] && [ -f gpurun_out/golden/ref_cuda_ops.npz ]; then
  cp gpurun_out/golden/ref_cuda_ops.npz tests/golden/ref_cuda_ops.npz
fi
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
timeout 300 python scripts/fps_sweep.py > gpurun_out/fps_sweep.log 2>&1
tail -5 gpurun_out/goldens.log gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.log
tail -40 gpurun_out/fps_sweep.log
