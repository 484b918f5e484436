#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 200 python bench.py --steps 50 --warmup 3 --lanes 8 --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e > gpurun_out/bench_tmp.log 2>&1
  python - "$@" <<'PY'
import json,sys,os
l=[x for x in open('gpurun_out/bench_tmp.log') if x.startswith('{')]
env={k:v for k,v in os.environ.items() if k.startswith('PN2_')}
if l:
    d=json.loads(l[-1]); print(env,'value %.0f scenes/s  %.3f ms/step'%(d['value'],d['ms_per_step']))
else: print(env, open('gpurun_out/bench_tmp.log').read()[-800:])
PY
}
run
PN2_SA_TC_V2=0 run
PN2_SA_TC_V2=0 PN2_SA_TC_PIPE=0 run
PN2_SA_TC_V2=0 PN2_SA_TC_PIPE=0 PN2_SA_SPLIT=0 run
PN2_SA_SPLIT=0 run
