timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sa_tc_v3 -s 2 -c 1 -o gpurun_out/sa1src python scripts/sa1_iso.py 2>&1 | tail -2
