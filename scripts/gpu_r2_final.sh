#!/bin/bash
# refresh of the committed evidence at HEAD: the driver's bench invocation, GPU tests, smoke, the launch list, and the
# full ncu capture of the throughput-mode sampling kernel (bucketed, 40 000 points) alone and with one scene per SM.
# (The full capture of one latency-mode step is scripts/gpu_profile_r2.sh; together they exceed the 64 MiB that travel
# back from the box in one call.)
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fps_bucket_kernel" -s 1 -c 1 -o gpurun_out/prof_fpsb40k_r2 -f python scripts/fps_bucket_one.py 40000 2048 8 > gpurun_out/ncu_fpsb40k.log 2>&1
tail -1 gpurun_out/ncu_fpsb40k.log
timeout 600 ncu --set full --clock-control none -k regex:"fps_bucket_kernel" -s 1 -c 1 -o gpurun_out/prof_fpsb40k_sat_r2 -f python scripts/fps_bucket_one.py 40000 2048 148 > gpurun_out/ncu_fpsb40k_sat.log 2>&1
tail -1 gpurun_out/ncu_fpsb40k_sat.log
B="python bench.py --warmup 3 --lanes 1 --no-graphs --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-e2e --no-sub-configs --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $B --steps 2 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log
ls -la gpurun_out
