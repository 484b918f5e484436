timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_fused_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python scripts/configs_bench.py 2>&1 | grep "re-encoding alone"
timeout 300 python bench.py --steps 20 --warmup 3 --no-sub-configs --no-reference-cuda --no-cpu-baseline --no-e2e > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -c 600 gpurun_out/bench_c.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c.json"))
print("value", d["value"], "ms", d["ms_per_step"], "parity", d["parity"]["ok"], "launches", d["gpu_launches_per_step"])
for r in d["roofline_kernels"]:
    if "ball" in r["kernel"]: print(r["kernel"], r["us"])
PY
