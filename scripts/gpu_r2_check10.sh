#!/bin/bash
# bucketed sampling kernel, two scenes per SM (PN2_FPS_BUCKET_SPLIT=2: 320 threads, histogram in the workspace):
# exactness (the whole FPS test set through it), one launch alone / saturated, and the bench step
mkdir -p gpurun_out
PN2_FPS_BUCKET_MIN=1 PN2_FPS_BUCKET_SPLIT=2 timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "furthest or fps" 2>&1 | tail -2
for sp in 0 2; do for b in 8 148 296; do echo -n "SPLIT=$sp B=$b: "; PN2_FPS_BUCKET_SPLIT=$sp python - <<PY
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
os.environ["PN2_FPS_BUCKET_MIN"] = "1"
from situation3d_b200._lib import check, lib, ptr, stream_ptr
from situation3d_b200.synthetic import make_scene
B, n, m = $b, 40000, 2048
base = np.stack([make_scene(s, n, 0)[:, :3] for s in range(8)])
xyz = torch.from_numpy(np.concatenate([base] * ((B + 7) // 8))[:B]).cuda().contiguous()
nb = lib.pn2_furthest_point_sampling_workspace_bytes(B, n, m)
ws = torch.empty(nb, dtype=torch.uint8, device="cuda"); idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
ts = []
for _ in range(4):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); check(lib.pn2_furthest_point_sampling_xyz_ws(B, n, m, ptr(xyz), ptr(idx), None, ptr(ws), nb, stream_ptr()), "fps"); e.record(); e.synchronize()
    ts.append(s.elapsed_time(e))
print("%.3f ms per launch, %.3f ms per 8 scenes" % (min(ts), min(ts) * 8 / B))
PY
done; done
run() { echo -n "$* $ARGS: "; env "$@" timeout 900 python bench.py --no-cpu-baseline --no-kernel-breakdown --no-reference-cuda --no-sub-configs --no-e2e $ARGS 2>gpurun_out/err.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['parity']['ok'], d['config'].get('sampling_mode'))" || tail -3 gpurun_out/err.txt; }
for rep in 1 2; do for sp in 0 2; do
ARGS="--steps 20 --warmup 5" run PN2_FPS_BUCKET_SPLIT=$sp
done; done
ARGS="--steps 20 --warmup 5 --lanes 32" run PN2_FPS_BUCKET_SPLIT=2
ARGS="--steps 20 --warmup 5 --lanes 40" run PN2_FPS_BUCKET_SPLIT=2
ARGS="--steps 64 --warmup 3 --lanes 40" run PN2_FPS_BUCKET_SPLIT=2
ARGS="--steps 64 --warmup 3" run PN2_FPS_BUCKET_SPLIT=2
