#!/bin/bash
# round-2: training step with the hand-written BatchNorm/ReLU/pool kernels and the prefetched sampling chain
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q -m gpu -k "rows_bn or row_layout or launch" 2>&1 | tail -5
timeout 600 python bench.py --config 4 --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_config4.json | cut -c1-1800
timeout 300 python scripts/train_profile.py rows 2>&1 | grep -v Warn | tee gpurun_out/train_profile_rows2.txt | head -45
