#!/bin/bash
# refresh of the committed evidence at HEAD: full default bench (the line the driver prints), then the ncu passes
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_full.json | cut -c1-1500
bash scripts/gpu_profile_r2.sh
