"""How the overlapped step time splits: the sampling chains alone vs everything else alone, 8 streams, graph replay (GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200 import fused
from situation3d_b200.backbone_module import Pointnet2Backbone
from situation3d_b200.synthetic import make_batch, randomize_bn_stats
torch.manual_seed(0)
net = randomize_bn_stats(Pointnet2Backbone(129, precision="bf16")).eval().cuda()
pcs = [torch.from_numpy(make_batch(8, 40000, 129, first_seed=8 * i)).cuda() for i in range(2)]
L, K = 8, 64
streams = [torch.cuda.Stream() for _ in range(L)]

def capture(fn):
    gs = []
    for i, st in enumerate(streams):
        with torch.cuda.stream(st):
            fn(i); fn(i)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            fn(i)
        gs.append(g)
    return gs

def run(gs, what):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    base = torch.cuda.current_stream()
    for rep in range(2):
        s.record(base)
        for st in streams: st.wait_event(s)
        for k in range(K):
            with torch.cuda.stream(streams[k % L]): gs[k % L].replay()
        for st in streams: base.wait_stream(st)
        e.record(base); torch.cuda.synchronize()
    print("%-42s %.3f ms/step" % (what, s.elapsed_time(e) / K))

with torch.no_grad():
    xyz = [pc[..., :3].contiguous() for pc in pcs]
    sas = (net.sa1, net.sa2, net.sa3, net.sa4)
    def fps_chain(i, levels=4):
        src = xyz[i % 2]
        for m in sas[:levels]:
            _, src = fused.fps_with_xyz(src, m.npoint)
    run(capture(lambda i: fps_chain(i, 1)), "FPS SA1 only")
    run(capture(lambda i: fps_chain(i, 4)), "FPS SA1-SA4 chain only")
    # everything else: precomputed sampling, then the main-stream kernels
    pre = []
    for i in range(2):
        src, lv = xyz[i], []
        for m in sas:
            inds, c = fused.fps_with_xyz(src, m.npoint); lv.append((inds, c)); src = c
        pre.append(lv)
    imgs = net._fused_images(pcs[0])
    def rest(i):
        pc = pcs[i % 2]; lv = pre[i % 2]
        src_xyz, table, ld, c, skip = xyz[i % 2], pc[..., 3:], 132, 129, 3
        rows = []
        for lvl, m in enumerate(sas):
            cx = lv[lvl][1]
            idx = fused.ball_query(src_xyz, cx, m.radius, m.nsample)
            out, out_rows = fused.sa_forward_bf16(imgs[lvl], src_xyz, cx, idx, table, ld, c, True, 1.0 / m.radius, raw_skip=skip)
            rows.append(out_rows)
            src_xyz, table, ld, c, skip = cx, out_rows, out_rows.shape[2], out_rows.shape[2], 0
        d2, i3 = fused.three_nn(lv[2][1], lv[3][1])
        _, f1 = fused.fp_forward_bf16(imgs[4], d2, i3, rows[3], rows[2])
        d2, i3 = fused.three_nn(lv[1][1], lv[2][1])
        fused.fp_forward_bf16(imgs[5], d2, i3, f1, rows[1], want_rows=False)
    run(capture(rest), "everything but the sampling chains")
    def sa_only(i):
        pc = pcs[i % 2]; lv = pre[i % 2]
        idx = sa_only.idx[i % 2]
        fused.sa_forward_bf16(imgs[0], xyz[i % 2], lv[0][1], idx, pc[..., 3:], 132, 129, True, 1.0 / sas[0].radius, raw_skip=3)
    sa_only.idx = [fused.ball_query(xyz[i], pre[i][0][1], sas[0].radius, sas[0].nsample) for i in range(2)]
    run(capture(sa_only), "SA1 (pointwise + fused) only")
    def bq_only(i):
        fused.ball_query(xyz[i % 2], pre[i % 2][0][1], sas[0].radius, sas[0].nsample)
    run(capture(bq_only), "ball query SA1 only")
