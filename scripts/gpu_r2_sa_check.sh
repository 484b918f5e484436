timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_fused_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 100 python scripts/sa1_iso.py 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 3 --no-sub-configs --no-reference-cuda --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -c 600 gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_b.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"]["ok"])
for r in d["roofline_kernels"]: print(r["kernel"], r["us"], r.get("frac"), r.get("layer_tflops"))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sa_tc_v3 -s 4 -c 1 -o gpurun_out/sa1src2 python scripts/sa1_iso.py 2>&1 | tail -1
