for v in 1 0 1 0; do
PN2_FP_TC2=$v timeout 300 python bench.py --steps 40 --warmup 3 --no-sub-configs --no-reference-cuda --no-cpu-baseline --no-e2e --no-parity > gpurun_out/bench_e$v.json 2> gpurun_out/bench_e.err; tail -c 300 gpurun_out/bench_e.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_e$v.json"))
print("PN2_FP_TC2=$v value", round(d["value"]), "ms", round(d["ms_per_step"],4), "launches", d["gpu_launches_per_step"], [ (r["kernel"], round(r["us"],1)) for r in d["roofline_kernels"] if "fp" in r["kernel"] or "three" in r["kernel"]])
PY
done
