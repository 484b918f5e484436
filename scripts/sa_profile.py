"""Per-phase cycle profile of the warp-specialised SA tensor-core kernel on the bench workload (GPU box)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200 import fused
from situation3d_b200._lib import check, lib, ptr
from situation3d_b200.backbone_module import Pointnet2Backbone
from situation3d_b200.synthetic import make_batch, randomize_bn_stats

torch.manual_seed(0)
net = randomize_bn_stats(Pointnet2Backbone(129, precision="bf16")).eval().cuda()
pc = torch.from_numpy(make_batch(8, 40000, 129)).cuda()
prof = torch.zeros(32, dtype=torch.int64, device="cuda")
V2 = os.environ.get("PN2_SA_TC_V2", "1") != "0"
with torch.no_grad():
    net({"point_clouds": pc})
    imgs = net._fused_images(pc)
    xyz = pc[..., :3].contiguous()
    src_xyz, table, ld, c = xyz, pc[..., 3:], 132, 129
    names = ["wait full", "mma1", "epi1+sync", "mma2", "epi2+sync", "mma3", "epi3+sync"]
    for lvl, m in enumerate((net.sa1, net.sa2, net.sa3, net.sa4)):
        inds, cxyz = fused.fps_with_xyz(src_xyz, m.npoint)
        idx = fused.ball_query(src_xyz, cxyz, m.radius, m.nsample)
        skip = 3 if lvl == 0 else 0
        run = lambda: fused.sa_forward_bf16(imgs[lvl], src_xyz, cxyz, idx, table, ld, c, True, 1.0 / m.radius, raw_skip=skip)
        run()
        check(lib.pn2_debug_sa_tc_profile(ptr(prof)), "prof")
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out, rows = run(); e.record(); e.synchronize()
        check(lib.pn2_debug_sa_tc_profile(None), "prof")
        pr = prof.cpu().numpy()
        tiles = max(int(pr[7]), 1)
        print("SA%d: %.1f us (pack / per-point GEMM + fused kernel), CTA0 tiles %d" % (lvl + 1, 1e3 * s.elapsed_time(e), tiles))
        if V2:
            f = lambda a, names: {n: int(v / tiles) for n, v in zip(names, a)}
            print("  epilogue cycles/tile:", f(pr[0:6], ["wait d1", "e1", "wait d2", "e2", "wait d3", "e3"]), "sum", int(pr[:6].sum() / tiles))
            print("  issuer A cycles/tile:", f(pr[8:11], ["wait full", "wait tfree", "issue L1"]))
            print("  issuers B, C cycles/tile:", f(pr[11:15], ["wait a1", "issue L2", "wait a2", "issue L3"]))
            print("  gather   cycles/tile:", f(pr[16:19], ["wait empty", "issue gathers", "xyz chunk"]))
        else:
            print("  consumer cycles/tile:", {n: int(v / tiles) for n, v in zip(names, pr[:7])}, "sum", int(pr[:7].sum() / tiles))
            print("  producer cycles/tile: wait empty %d, issue gathers %d, xyz chunk %d" % tuple(int(v / tiles) for v in pr[8:11]))
        src_xyz, table, ld, c = cxyz, rows, rows.shape[2], rows.shape[2]
