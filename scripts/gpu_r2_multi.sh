timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n2.json"))
print("N=2 value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["ok"], d["parity"]["all_ranks_ok"])
for k,v in d.get("configs",{}).items():
    for r in v: print(k, r.get("points"), round(r.get("value",0),1), round(r.get("ms_per_step",0),3), (r.get("parity") or {}).get("ok"), r.get("allreduce_us_alone"), r.get("error"))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
