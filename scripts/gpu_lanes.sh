#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 300 python bench.py --steps 48 --warmup 3 --lanes $1 --fps-sms $2 --no-cpu-baseline $3 --no-reference-cuda > gpurun_out/bench_tmp.log 2>&1
  python - "$@" <<'PY'
import json,sys
l=[x for x in open('gpurun_out/bench_tmp.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('lanes %s fps_sms %s -> %s'%(sys.argv[1],sys.argv[2],d['config'].get('sm_partition')),'value %.0f scenes/s  %.3f ms/step | e2e %.0f  %.2f ms/step'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))
    for r in d.get('roofline_kernels',[]):
        if r['kernel'].startswith('fps'): print('   ',r['kernel'], round(r['us'],1),'us', '%.2f M rounds/s'%(r['rounds_per_s']/1e6))
else: print(open('gpurun_out/bench_tmp.log').read()[-1500:])
PY
}
run 1 0; run 6 0 --no-kernel-breakdown; run 8 0 --no-kernel-breakdown; run 8 80 --no-kernel-breakdown; run 12 96 --no-kernel-breakdown
