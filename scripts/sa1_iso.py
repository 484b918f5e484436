"""SA1 fused kernel alone (P rows), device time under the diagnostic switches (GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200 import fused
from situation3d_b200._lib import check, lib, ptr, stream_ptr
from situation3d_b200.backbone_module import Pointnet2Backbone
from situation3d_b200.synthetic import make_batch, randomize_bn_stats
torch.manual_seed(0)
net = randomize_bn_stats(Pointnet2Backbone(129, precision="bf16")).eval().cuda()
pc = torch.from_numpy(make_batch(8, 40000, 129)).cuda()
flush = torch.empty(1024 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
def timeit(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]
with torch.no_grad():
    net({"point_clouds": pc})
    imgs = net._fused_images(pc)
    B, N, W = pc.shape
    xyz = pc[..., :3].contiguous()
    img = imgs[0]
    lin_image, sa_image = img.split(132, 3, False)
    P = fused.lin_rows(lin_image, pc, ptr(pc), B * N, 132, 64, False, 132).view(B, N, 64)
    m = net.sa1
    inds, cxyz = fused.fps_with_xyz(xyz, m.npoint)
    idx = fused.ball_query(xyz, cxyz, m.radius, m.nsample)
    out = torch.empty((B, 128, m.npoint), device="cuda"); rows = torch.empty((B, m.npoint, 128), dtype=torch.bfloat16, device="cuda")
    t = timeit(lambda: check(lib.pn2_sa_tc_forward(B, N, m.npoint, m.nsample, 64, 64, 64, 128, 1.0 / m.radius, ptr(xyz), ptr(cxyz),
                                                   ptr(P), ptr(idx), ptr(sa_image), ptr(out), ptr(rows), stream_ptr()), "sa"))
    print("SA1 fused over P rows [%s]: %.1f us" % (" ".join("%s=%s" % kv for kv in os.environ.items() if kv[0].startswith("PN2_")), t))
