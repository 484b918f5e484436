#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -20
grep -B5 -A25 "Error\|assert " gpurun_out/pytest_gpu.log | head -60
for PIPE in 0 1; do
export PN2_SA_TC_PIPE=$PIPE
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-cuda > gpurun_out/bench_pipe$PIPE.log 2>&1
python - $PIPE <<'PY'
import json,sys
l=[x for x in open('gpurun_out/bench_pipe%s.log'%sys.argv[1]) if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print('PIPE',sys.argv[1],d['value'], d['ms_per_step'], d['e2e']['value'])
    for r in d.get('roofline_kernels', []):
        if 'fused' in r['kernel']: print("%-22s %9.1f us  share %.3f  %s %.3f %s frac %.4f" % (r['kernel'], r['us'], r['share'], r['bound'], r['achieved'], r['unit'], r['frac']))
else:
    print(open('gpurun_out/bench_pipe%s.log'%sys.argv[1]).read()[-2000:])
PY
done
