"""Per-phase cycles of one FPS round (thread 0 of CTA 0), bench shapes (GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200._lib import check, lib, ptr, stream_ptr
from situation3d_b200.synthetic import make_scene
B = 8
xyz = torch.from_numpy(np.stack([make_scene(s, 40000, 0)[:, :3] for s in range(B)])).cuda().contiguous()
print("n->m: cycles/round [update+candidate, warp argmax, exchange, table reduce] | us/round")
for n, m in [(40000, 2048), (2048, 1024), (1024, 512), (512, 256)]:
    src = xyz[:, :n].contiguous()
    idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
    prof = torch.zeros(8, dtype=torch.int64, device="cuda")
    for _ in range(2):
        check(lib.pn2_debug_fps_profile(B, n, m, ptr(src), ptr(idx), ptr(prof), stream_ptr()), "prof")
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); check(lib.pn2_debug_fps_profile(B, n, m, ptr(src), ptr(idx), ptr(prof), stream_ptr()), "prof"); e.record(); e.synchronize()
    p = prof.cpu().numpy()[:4] / (m - 1)
    print("%d->%d: %s sum %.0f | %.3f us/round" % (n, m, np.round(p, 0), p.sum(), 1e3 * s.elapsed_time(e) / (m - 1)))
