#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -20
grep -B5 -A25 "Error\|assert" gpurun_out/pytest_gpu.log | head -80
run() {
  timeout 300 python bench.py --steps 20 --warmup 3 --lanes $1 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda > gpurun_out/bench_tmp.log 2>&1
  python - "$1" "$2" <<'PY'
import json,sys
l=[x for x in open('gpurun_out/bench_tmp.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('lanes',sys.argv[1],sys.argv[2],'value %.0f scenes/s  %.2f ms/step | e2e %.0f  %.2f ms/step'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))
else: print(open('gpurun_out/bench_tmp.log').read()[-1500:])
PY
}
run 1 default; run 4 default; run 6 default; run 8 default; run 12 default
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
