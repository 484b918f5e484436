"""Sweep the FPS cluster shape (PN2_FPS_CLUSTER x PN2_FPS_THREADS) on the bench workload and the
smaller pyramid levels; prints ms per launch and rounds/s.  GPU box only."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200 import fused  # noqa: E402
from situation3d_b200.synthetic import make_scene  # noqa: E402


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def main():
    B = int(os.environ.get("B", "8"))
    xyz = torch.from_numpy(np.stack([make_scene(s, 40000, 0)[:, :3] for s in range(B)])).cuda().contiguous()
    print("level n->m, cluster, threads, ms, us/round")
    for n, m in [(40000, 2048), (2048, 1024), (1024, 512), (512, 256)]:
        src = xyz[:, :n].contiguous()
        for cl in (1, 2, 4, 8, 16):
            for th in (128, 256, 512):
                per = -(-n // (cl * th))
                if per > 32 or (n <= 2048 and cl > 2) or (n > 8192 and cl * th < 2048):
                    continue
                os.environ["PN2_FPS_CLUSTER"], os.environ["PN2_FPS_THREADS"] = str(cl), str(th)
                try:
                    ms = timeit(lambda: fused.fps_with_xyz(src, m))
                    print("%d->%d, %d, %d, %.3f, %.3f" % (n, m, cl, th, ms, 1e3 * ms / (m - 1)), flush=True)
                except Exception as ex:
                    print("%d->%d, %d, %d, failed: %s" % (n, m, cl, th, ex), flush=True)
    os.environ.pop("PN2_FPS_CLUSTER"), os.environ.pop("PN2_FPS_THREADS")


def profile():
    from situation3d_b200._lib import check, lib, ptr, stream_ptr
    B = 8
    xyz = torch.from_numpy(np.stack([make_scene(s, 40000, 0)[:, :3] for s in range(B)])).cuda().contiguous()
    print("phase profile (cycles/round of thread 0, CTA 0): update, warp argmax, exchange, scene argmax | wall us/round -> implied MHz")
    for n, m in [(40000, 2048), (2048, 1024), (512, 256)]:
        src = xyz[:, :n].contiguous()
        idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
        prof = torch.zeros(8, dtype=torch.int64, device="cuda")
        for _ in range(2):
            check(lib.pn2_debug_fps_profile(B, n, m, ptr(src), ptr(idx), ptr(prof), stream_ptr()), "prof")
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        check(lib.pn2_debug_fps_profile(B, n, m, ptr(src), ptr(idx), ptr(prof), stream_ptr()), "prof")
        e.record()
        e.synchronize()
        us = 1e3 * s.elapsed_time(e) / (m - 1)
        pc = prof.cpu().numpy()[:5] / (m - 1)
        print("%d->%d: %s | %.3f us/round -> %.0f MHz" % (n, m, np.round(pc, 1), us, pc[:4].sum() / us))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "profile":
        profile()
    else:
        main()
