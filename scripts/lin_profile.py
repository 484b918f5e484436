"""Device time of the per-point GEMM (lin_tc) and of the fused SA kernel alone, bench workload (GPU box)."""
import ctypes, os, sys
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
        flush.zero_()      # 1 GB: evicts L2 and keeps the GPU busy while the host enqueues fn
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
    P = torch.empty((B * N, 64), dtype=torch.bfloat16, device="cuda")
    t = timeit(lambda: check(lib.pn2_lin_tc_forward(B * N, 132, 64, ptr(pc), 0, 132, ptr(lin_image), ptr(P), stream_ptr()), "lin"))
    print("lin SA1 (320000 rows x 132 f32 -> 64 bf16): %.1f us  = %.0f GB/s" % (t, (pc.numel() * 4 + P.numel() * 2) / t / 1e3))
    tab = torch.empty((B, N, 136), dtype=torch.bfloat16, device="cuda")
    t = timeit(lambda: check(lib.pn2_sa_tc_pack_rows(B, N, 129, ptr(pc[..., 3:]), N * 132, 132, ptr(tab), stream_ptr()), "pack"))
    print("pack_rows SA1: %.1f us = %.0f GB/s" % (t, (pc.numel() * 4 + tab.numel() * 2) / t / 1e3))
    m = net.sa1
    inds, cxyz = fused.fps_with_xyz(xyz, m.npoint)
    idx = fused.ball_query(xyz, cxyz, m.radius, m.nsample)
    out = torch.empty((B, 128, m.npoint), device="cuda"); rows = torch.empty((B, m.npoint, 128), dtype=torch.bfloat16, device="cuda")
    Pv = P.view(B, N, 64)
    t = timeit(lambda: check(lib.pn2_sa_tc_forward(B, N, m.npoint, m.nsample, 64, 64, 64, 128, 1.0 / m.radius, ptr(xyz), ptr(cxyz),
                                                   ptr(Pv), ptr(idx), ptr(sa_image), ptr(out), ptr(rows), stream_ptr()), "sa"))
    print("SA1 fused over P rows: %.1f us" % t)
    t = timeit(lambda: check(lib.pn2_sa_tc_forward(B, N, m.npoint, m.nsample, 129, 64, 64, 128, 1.0 / m.radius, ptr(xyz), ptr(cxyz),
                                                   ptr(tab), ptr(idx), ptr(img.image), ptr(out), ptr(rows), stream_ptr()), "sa"))
    print("SA1 fused over feature rows: %.1f us" % t)
    # SA3-shaped bf16 lin
    r3 = torch.randn(8 * 1024, 256, device="cuda").bfloat16()
    li3, _ = imgs[2].split(256, 0, True)
    P3 = torch.empty((8 * 1024, 128), dtype=torch.bfloat16, device="cuda")
    t = timeit(lambda: check(lib.pn2_lin_tc_forward(8 * 1024, 256, 128, ptr(r3), 1, 256, ptr(li3), ptr(P3), stream_ptr()), "lin"))
    print("lin SA3 (8192 rows x 256 bf16 -> 128): %.1f us" % t)
