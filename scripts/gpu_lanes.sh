#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_gpu.py -m gpu -q --timeout 200 -x -k "graphed or partition" 2>&1 | tail -4
run() {
  timeout 300 python bench.py --steps $1 --warmup 3 --lanes $2 $3 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda > gpurun_out/bench_tmp.log 2>&1
  python - "$@" <<'PY'
import json,sys,os
l=[x for x in open('gpurun_out/bench_tmp.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('steps %s lanes %s %s'%(sys.argv[1],sys.argv[2],sys.argv[3:]),'value %.0f scenes/s  %.3f ms/step, host enqueue %.3f ms/step | e2e %.0f'%(d['value'],d['ms_per_step'],d['host_enqueue_ms_per_step'],d['e2e']['value']))
else: print(open('gpurun_out/bench_tmp.log').read()[-2500:])
PY
}
run 50 8 --no-graphs; run 50 1; run 50 4; run 50 6; run 50 8; run 50 12; run 20 8; run 100 8
