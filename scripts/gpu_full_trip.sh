#!/bin/bash
# full GPU parity suite + smoke + bench (what the driver runs at round end)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|exit" gpurun_out/pytest_gpu.log | head -20
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py --steps 30 --warmup 3 $BENCH_ARGS > gpurun_out/bench_full.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_full.log
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_full.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('value %.0f scenes/s  %.2f ms/step | e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']))
    for r in d.get('roofline_kernels',[]): print(r['kernel'], round(r['us'],1),'us', round(r['frac'],3))
    print({k: d.get(k) for k in ('cpu_baseline','reference_cuda','clocks','gpu_launches')})
else: print(open('gpurun_out/bench_full.log').read()[-1500:])
PY
