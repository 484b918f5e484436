timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 --no-sub-configs --no-reference-cuda --no-cpu-baseline --no-e2e > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -c 600 gpurun_out/bench_d.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_d.json"))
print("value", d["value"], "ms", d["ms_per_step"], "parity", d["parity"]["ok"], "launches", d["gpu_launches_per_step"])
for r in d["roofline_kernels"]:
    if "fp" in r["kernel"] or "three" in r["kernel"]: print(r["kernel"], r["us"], r.get("frac"))
PY
timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_fused_gpu.py -x -q -m gpu -k "fp_module" 2>&1 | tail -4
