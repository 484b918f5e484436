#!/bin/bash
# last check at HEAD: full GPU test-suite, smoke, the driver's bench invocation
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_full.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['roofline']['kernel'], d['roofline']['traffic'], [ (k, round(v[0]['value'])) for k,v in d['configs'].items()])"
