#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -20
for L in 1 2 3 4 6; do
  timeout 300 python bench.py --steps 20 --warmup 3 --lanes $L --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda > gpurun_out/bench_l$L.log 2>&1
  python - $L <<'PY'
import json,sys
L=sys.argv[1]
l=[x for x in open('gpurun_out/bench_l%s.log'%L) if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('lanes',L,'value %.0f scenes/s  %.2f ms/step | e2e %.0f  %.2f ms/step'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']), d['clocks'])
else: print(open('gpurun_out/bench_l%s.log'%L).read()[-1500:])
PY
done
timeout 300 python scripts/fps_sweep.py profile > gpurun_out/fps_profile.log 2>&1; cat gpurun_out/fps_profile.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_bf16.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_bf16.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'])
    for r in d.get('roofline_kernels', []):
        print("%-22s %9.1f us  share %.3f  %s %.3f %s frac %.4f" % (r['kernel'], r['us'], r['share'], r['bound'], r['achieved'], r['unit'], r['frac']))
else:
    print(open('gpurun_out/bench_bf16.log').read()[-2000:])
PY
