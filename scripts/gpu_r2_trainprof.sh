#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q -m gpu -k "rows_bn or row_layout or group_rows" 2>&1 | tail -2
timeout 300 python scripts/train_prefetch_diag.py 2>&1 | grep prefetch | cut -c1-150
timeout 600 python bench.py --config 4 --steps 24 2>&1 | tail -1 | tee gpurun_out/bench_config4.json | python -c "import sys,json; t=json.loads(sys.stdin.read())['train']; print({k:t[k] for k in ['ms_per_step','ms_per_step_median','ms_per_step_without_sampling_prefetch','reference_wiring_ms_per_step']}, t['parity']['ok'])"
