"""The fused FP layers alone at the bench shapes (B = 8): device time per launch, for ncu (GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200 import fused
from situation3d_b200.backbone_module import Pointnet2Backbone
from situation3d_b200.synthetic import make_batch, randomize_bn_stats
os.environ.setdefault("PN2_FP_TC2", "1")      # the cluster kernel (at B = 8 the library would pick the single-CTA kernel)
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
    out = net({"point_clouds": pc})
    imgs = net._fused_images(pc)
    B = 8
    g = torch.Generator(device="cuda").manual_seed(1)
    rows = {n: torch.randn(B, n, 256, device="cuda", generator=g).relu().bfloat16() for n in (256, 512, 1024)}
    for name, img, un, kn in (("fp1", imgs[4], out["sa3_xyz"], out["sa4_xyz"]), ("fp2", imgs[5], out["sa2_xyz"], out["sa3_xyz"])):
        n, m = un.shape[1], kn.shape[1]
        t = timeit(lambda: fused.fp_layer("bf16", img, un, kn, rows[m], rows[n]))
        from situation3d_b200._lib import lib, ptr
        prof = torch.zeros(16, dtype=torch.int64, device="cuda")
        lib.pn2_debug_fp_tc2_profile(ptr(prof)); fused.fp_layer("bf16", img, un, kn, rows[m], rows[n]); torch.cuda.synchronize()
        lib.pn2_debug_fp_tc2_profile(None)
        pr = prof.cpu().numpy()
        print("   phase cycles (setup, three_nn, exchange, A operand, layer 1, barrier, epilogue 1, barrier, layer 2, epilogue 2):",
              [int(b - a) for a, b in zip(pr[:10], pr[1:11])])
        print("%s layer (n=%d, m=%d) [%s]: %.1f us" % (name, n, m, " ".join("%s=%s" % kv for kv in os.environ.items() if kv[0].startswith("PN2_")), t))
