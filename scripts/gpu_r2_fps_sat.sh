#!/bin/bash
mkdir -p gpurun_out
export PN2_FPS_BUCKET_MIN=1000000000
for b in 8 16 32 37 48 64; do timeout 120 python scripts/fps_sat_one.py $b | tail -1; done
timeout 600 ncu --set full --import-source on --clock-control none -s 1 -c 1 -o gpurun_out/fps_sat32 -f python scripts/fps_sat_one.py 32 > gpurun_out/fps_sat32.log 2>&1
tail -3 gpurun_out/fps_sat32.log
timeout 600 ncu --set full --import-source on --clock-control none -s 1 -c 1 -o gpurun_out/fps_sat8 -f python scripts/fps_sat_one.py 8 > gpurun_out/fps_sat8.log 2>&1
tail -2 gpurun_out/fps_sat8.log
