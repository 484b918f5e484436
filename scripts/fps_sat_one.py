"""One FPS launch that fills the GPU with two CTAs per SM (B = 32 scenes x 8-CTA clusters = 256 CTAs), for ncu:
the co-resident steady state a pipelined step runs in, which a serialised capture of the step cannot show.
    ncu --set full -s 1 -c 1 python scripts/fps_sat_one.py [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200._lib import check, lib, ptr, stream_ptr
from situation3d_b200.synthetic import make_scene
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n, m = 40000, 2048
base = np.stack([make_scene(s, n, 0)[:, :3] for s in range(8)])
xyz = torch.from_numpy(np.concatenate([base] * ((B + 7) // 8))[:B]).cuda().contiguous()
idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
nx = torch.empty((B, m, 3), dtype=torch.float32, device="cuda")
for _ in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(lib.pn2_furthest_point_sampling_xyz(B, n, m, ptr(xyz), ptr(idx), ptr(nx), stream_ptr()), "fps")
    e.record(); e.synchronize()
    print("B=%d: %.3f ms per launch, %.3f ms per 8 scenes" % (B, s.elapsed_time(e), s.elapsed_time(e) * 8 / B))
