#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 200 python bench.py --steps $1 --warmup 3 --lanes $2 $3 $4 $5 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e > gpurun_out/bench_tmp.log 2>&1
  python - "$@" <<'PY'
import json,sys,os
l=[x for x in open('gpurun_out/bench_tmp.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('steps %s lanes %s %s'%(sys.argv[1],sys.argv[2],sys.argv[3:]),'value %.0f scenes/s  %.3f ms/step, host enqueue %.3f ms/step'%(d['value'],d['ms_per_step'],d['host_enqueue_ms_per_step']), d['config'].get('sm_partition'))
else: print(open('gpurun_out/bench_tmp.log').read()[-1200:])
PY
}
run 50 8 --fps-sms 80; run 50 8 --fps-sms 96; run 50 12 --fps-sms 88; run 50 8 --fps-sms 64
export PN2_FPS_MINB=1
run 50 8
