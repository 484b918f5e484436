timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 100 python scripts/sa1_iso.py 2>&1 | tail -1
timeout 200 python scripts/configs_bench.py 2>&1 | grep "re-encoding alone" | cut -c1-260
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fps_kernel|sa_tc_v3" -s 8 -c 8 python scripts/sa1_iso.py 2>&1 | grep -E "fps_kernel|sa_tc_v3|dram__bytes|gpu__time" | head -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
