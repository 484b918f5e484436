#!/bin/bash
# pipelined SA kernel + split first layer: parity tests first (short timeouts: a hang must not hold the box), then profile + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_fused_gpu.py -m gpu -q --timeout 120 -x > gpurun_out/pytest_v2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_v2.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/pytest_v2.log | head -20
grep -B5 -A25 "Error\|assert " gpurun_out/pytest_v2.log | head -60
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 120 python scripts/lin_profile.py 2>&1 | tail -6
timeout 120 python scripts/sa_profile.py > gpurun_out/sa_profile_v2.log 2>&1; cat gpurun_out/sa_profile_v2.log
if [ "$1" == "bench" ]; then
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/bench_v2.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_v2.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('value %.0f scenes/s  %.2f ms/step | e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']))
    for r in d.get('roofline_kernels',[]): print(r['kernel'], round(r['us'],1),'us', round(r['frac'],3))
else: print(open('gpurun_out/bench_v2.log').read()[-1500:])
PY
fi
