"""Cluster FPS kernel, 40 slots per thread: how many slot pairs keep their coordinates in registers (PN2_FPS_RP).
Fewer register-resident pairs -> 168 registers -> three CTAs per SM instead of two.  Latency of one 8-scene launch
and saturated throughput with 8 streams in flight, every variant checked against the all-register kernel (GPU box).
    python scripts/fps_rp_sweep.py            # sweeps the variants, one subprocess each
"""
import hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    import numpy as np, torch
    sys.path.insert(0, ROOT)
    from situation3d_b200._lib import check, lib, ptr, stream_ptr
    from situation3d_b200.synthetic import make_scene
    n, m, B = 40000, 2048, 8
    xyz = torch.from_numpy(np.stack([make_scene(s, n, 0)[:, :3] for s in range(B)])).cuda().contiguous()

    def run(idx, nx):
        check(lib.pn2_furthest_point_sampling_xyz(B, n, m, ptr(xyz), ptr(idx), ptr(nx), stream_ptr()), "fps")

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); e.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.median(ts))

    idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
    nx = torch.empty((B, m, 3), dtype=torch.float32, device="cuda")
    out = {"rp": os.environ.get("PN2_FPS_RP", "20"), "ms_per_launch": timed(lambda: run(idx, nx))}
    out["sha"] = hashlib.sha256(idx.cpu().numpy().tobytes() + nx.cpu().numpy().tobytes()).hexdigest()[:16]
    for nstreams in (8, 12):
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
        bufs = [(torch.empty_like(idx), torch.empty_like(nx)) for _ in streams]

        def sat():
            for st, (i2, x2) in zip(streams, bufs):
                st.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(st):
                    for _ in range(4):
                        run(i2, x2)
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)
        out["saturated_ms_per_batch_%d_streams" % nstreams] = timed(sat, reps=3) / (4 * nstreams)
    print(json.dumps(out))
else:
    for rp in ("20", "6", "0"):
        env = dict(os.environ, PN2_FPS_RP=rp, PN2_FPS_BUCKET_MIN="1000000000")
        r = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or r.stderr[-400:], flush=True)
