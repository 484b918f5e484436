#!/bin/bash
mkdir -p gpurun_out
export PN2_FPS_BUCKET_MIN=1000000000
for rp in 20 6 4 0; do for b in 8 32 48; do echo -n "RP=$rp "; PN2_FPS_RP=$rp timeout 120 python scripts/fps_sat_one.py $b | tail -1; done; done
PN2_FPS_RP=6 timeout 600 ncu --set full --import-source on --clock-control none -s 1 -c 1 -o gpurun_out/fps_sat48_rp6 -f python scripts/fps_sat_one.py 48 > gpurun_out/fps_sat48.log 2>&1
tail -2 gpurun_out/fps_sat48.log
