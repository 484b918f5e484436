"""Bucketed FPS kernel (csrc/fps_bucket.cu) against the cluster kernel (csrc/fps.cu): latency of one launch,
per-phase cycles, and saturated throughput with several batches in flight (GPU box).
    python scripts/fps_bucket_bench.py [n] [m]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200._lib import check, lib, ptr, stream_ptr
from situation3d_b200.synthetic import make_scene

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
B = 8
xyz = torch.from_numpy(np.stack([make_scene(s, n, 0)[:, :3] for s in range(B)])).cuda().contiguous()
out = {"n": n, "m": m, "B": B}


def run(bucket, stream=None, idx=None, new_xyz=None, ws=None):
    os.environ["PN2_FPS_BUCKET_MIN"] = "1" if bucket else "1000000000"
    nbytes = lib.pn2_furthest_point_sampling_workspace_bytes(B, n, m)
    if ws is None and nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    check(lib.pn2_furthest_point_sampling_xyz_ws(B, n, m, ptr(xyz), ptr(idx), ptr(new_xyz), ptr(ws), nbytes, stream_ptr()), "fps")
    return ws


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


res = {}
for name, bucket in (("cluster", False), ("bucket", True)):
    idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
    nx = torch.empty((B, m, 3), dtype=torch.float32, device="cuda")
    ws = run(bucket, idx=idx, new_xyz=nx)
    res[name] = idx.clone()
    t = timed(lambda: run(bucket, idx=idx, new_xyz=nx, ws=ws))
    out[name + "_ms_per_launch"] = t
    out[name + "_us_per_round"] = 1e3 * t / (m - 1)
    # saturated: 8 streams x 4 launches each
    streams = [torch.cuda.Stream() for _ in range(8)]
    bufs = [(torch.empty_like(idx), torch.empty_like(nx), torch.empty_like(ws) if ws is not None else None) for _ in streams]
    def sat():
        for st, (i2, x2, w2) in zip(streams, bufs):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                for _ in range(4):
                    run(bucket, idx=i2, new_xyz=x2, ws=w2)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
    ts = timed(sat, reps=3)
    out[name + "_saturated_ms_per_batch"] = ts / 32
out["identical"] = bool(torch.equal(res["cluster"], res["bucket"]))

os.environ["PN2_FPS_BUCKET_MIN"] = "1"
nbytes = lib.pn2_furthest_point_sampling_workspace_bytes(B, n, m)
ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
prof = torch.zeros(8, dtype=torch.int64, device="cuda")
for _ in range(2):
    check(lib.pn2_debug_fps_bucket_profile(B, n, m, ptr(xyz), ptr(idx), ptr(prof), ptr(ws), nbytes, stream_ptr()), "prof")
torch.cuda.synchronize()
pr = prof.cpu().numpy()
p = pr[:6] / (m - m // 2)          # the kernel accumulates over rounds m/2 .. m-1 (steady state)
out["bucket_cycles_per_round"] = dict(zip(["tests+publish", "barrier1+scan", "updates+own_argmax", "barrier2", "final_argmax+refresh"], [float(round(v, 1)) for v in p[:5]]))
out["bucket_cycles_per_round"]["sum"] = float(round(p.sum(), 1))
out["bucket_sources_per_warp_round"] = float(pr[6]) / (m - m // 2)
# prologue cost: a launch with m = 1 does the binning only
idx1 = torch.empty((B, 1), dtype=torch.int32, device="cuda")
out["bucket_prologue_ms"] = timed(lambda: check(lib.pn2_furthest_point_sampling_xyz_ws(B, n, 1, ptr(xyz), ptr(idx1), None, ptr(ws), nbytes, stream_ptr()), "fps"))
print(json.dumps(out))
