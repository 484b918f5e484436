#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/sa_profile.py > gpurun_out/sa_profile.log 2>&1; cat gpurun_out/sa_profile.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -20
grep -B5 -A25 "Error\|assert " gpurun_out/pytest_gpu.log | head -60
run() {
  timeout 300 python bench.py --steps 30 --warmup 3 --lanes $1 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda > gpurun_out/bench_tmp.log 2>&1
  python - "$1" "$2" <<'PY'
import json,sys
l=[x for x in open('gpurun_out/bench_tmp.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('lanes',sys.argv[1],sys.argv[2],'value %.0f scenes/s  %.2f ms/step | e2e %.0f  %.2f ms/step'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))
else: print(open('gpurun_out/bench_tmp.log').read()[-1500:])
PY
}
run 6 minb1; run 8 minb1; run 6 minb1-again
export PN2_FPS_MINB=2
run 6 minb2; run 8 minb2; run 12 minb2
