timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 100 python scripts/fp_iso.py 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -c 800 gpurun_out/bench_full.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_full.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "e2e_f32", round(d["e2e_f32_input"]["value"]), "parity", d["parity"]["ok"], d["parity"].get("e2e_result_ok"))
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["share_of_step"])
print("step", d.get("roofline_step"))
print("refcuda", json.dumps(d.get("reference_cuda"))[:900])
print("cpu", d["cpu_baseline"]["value"])
for k,v in d["configs"].items():
    for r in v: print(k, r.get("points"), round(r.get("value",0),1), round(r.get("ms_per_step",0),3), (r.get("parity") or {}).get("ok"), r.get("error"))
for r in d["roofline_kernels"]: print(r["kernel"], round(r["us"],1), round(r["share"],3), round(r["frac"],4), r.get("layer_tflops"))
PY
