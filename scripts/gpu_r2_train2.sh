#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q -m gpu -k "rows_bn or row_layout or group_rows" 2>&1 | tail -3
timeout 600 python bench.py --config 4 --steps 12 2>&1 | tail -1 | tee gpurun_out/bench_config4.json | python -c "import sys,json; t=json.loads(sys.stdin.read())['train']; print({k:t[k] for k in ['ms_per_step','ms_per_step_without_sampling_prefetch','reference_wiring_ms_per_step']}, t['parity']['ok'])"
timeout 300 python scripts/train_profile.py rows 2>&1 | grep -E "ms per step|finalize|sum of kernel" 
