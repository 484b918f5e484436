#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q -m gpu -k "rows_bn or row_layout" 2>&1 | tail -3
echo "== fused BN kernels + split layer 1"; timeout 300 python scripts/train_parity_diag.py 2>&1 | grep -v Warn | grep -v detach | head -6
timeout 600 python bench.py --config 4 --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_config4.json | cut -c1-1500
PN2_TRAIN_SPLIT_L1=0 timeout 600 python bench.py --config 4 --steps 10 --no-reference-cuda 2>&1 | tail -1 | cut -c1-400
timeout 300 python scripts/train_profile.py rows 2>&1 | grep -v Warn | tee gpurun_out/train_profile_rows2.txt | head -36
