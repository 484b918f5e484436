timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_fused_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 100 python scripts/sa1_iso.py 2>&1 | tail -1
PN2_SA_TC_TMA=0 timeout 100 python scripts/sa1_iso.py 2>&1 | tail -1
for v in 1 0; do
PN2_SA_TC_TMA=$v timeout 300 python bench.py --steps 40 --warmup 3 --no-sub-configs --no-reference-cuda --no-cpu-baseline --no-e2e > gpurun_out/bench_t$v.json 2> gpurun_out/bench_t.err; tail -c 300 gpurun_out/bench_t.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_t$v.json"))
print("PN2_SA_TC_TMA=$v value", round(d["value"]), "ms", round(d["ms_per_step"],4), "parity", d["parity"]["ok"], [ (r["kernel"], round(r["us"],1)) for r in d["roofline_kernels"] if "fused" in r["kernel"]])
PY
done
