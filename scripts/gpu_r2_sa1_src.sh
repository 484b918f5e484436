timeout 300 ncu --set full --clock-control none --import-source on -k regex:sa_tc_v3 -s 4 -c 1 -o gpurun_out/sa1src python scripts/sa1_iso.py 2>&1 | tail -2
timeout 200 python -m pytest tests/test_fused_gpu.py -x -q -m gpu -k "reencode or smoke or situat" 2>&1 | tail -3
timeout 200 python scripts/configs_bench.py 2>&1 | grep "re-encoding alone"
