"""Two launches of the bucketed FPS kernel at the bench shape (for ncu): python scripts/fps_bucket_one.py [n] [m] [B]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200._lib import check, lib, ptr, stream_ptr
from situation3d_b200.synthetic import make_scene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
os.environ["PN2_FPS_BUCKET_MIN"] = "1"
xyz = torch.from_numpy(np.stack([make_scene(s, n, 0)[:, :3] for s in range(B)])).cuda().contiguous()
nbytes = lib.pn2_furthest_point_sampling_workspace_bytes(B, n, m)
ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
idx = torch.empty((B, m), dtype=torch.int32, device="cuda")
for _ in range(2):
    check(lib.pn2_furthest_point_sampling_xyz_ws(B, n, m, ptr(xyz), ptr(idx), None, ptr(ws), nbytes, stream_ptr()), "fps")
torch.cuda.synchronize()
print("ok", int(idx[0, 1]))
