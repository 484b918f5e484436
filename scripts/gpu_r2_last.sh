#!/bin/bash
# last check at HEAD: full GPU test-suite, smoke, the no-flag bench invocation
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_default.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['steps'], d['warmup'], d['e2e']['value'], d['parity']['ok'], d['roofline']['kernel'], d['roofline'].get('variant'), d['roofline']['traffic'])"
