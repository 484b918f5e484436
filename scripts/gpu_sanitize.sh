#!/bin/bash
# compute-sanitizer on the tensor-core and sampling kernels at small shapes (memcheck, then racecheck on shared memory)
mkdir -p gpurun_out
T="tests/test_tc_gpu.py -k (sa_tc_vs_bf16_emulation or lin_tc_row_gemm)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_tc_gpu.py -m gpu -q -x --timeout 600 -k "sa_tc_vs_bf16_emulation or lin_tc_row_gemm" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log
tail -4 gpurun_out/sanitize_memcheck.log; grep -c "Invalid\|Error:" gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --timeout 600 -k "fps or furthest" > gpurun_out/sanitize_memcheck_fps.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck_fps.log
tail -3 gpurun_out/sanitize_memcheck_fps.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_tc_gpu.py -m gpu -q -x --timeout 800 -k "lin_tc_row_gemm" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitize_racecheck.log
tail -4 gpurun_out/sanitize_racecheck.log
