#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/fps_sweep.py profile > gpurun_out/fps_profile.log 2>&1
timeout 300 python scripts/fps_sweep.py > gpurun_out/fps_sweep.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_bf16.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -40
cat gpurun_out/fps_profile.log; cat gpurun_out/fps_sweep.log
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
