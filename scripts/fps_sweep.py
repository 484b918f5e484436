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


if __name__ == "__main__":
    main()
