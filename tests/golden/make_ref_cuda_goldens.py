"""Golden vectors from the REFERENCE's own CUDA kernels (runs on the GPU box only).

    gpurun -- python tests/golden/make_ref_cuda_goldens.py      # writes gpurun_out/golden/ref_cuda_ops.npz

The extension in oracle/_ref/ is the reference's lib/pointnet2/_ext_src compiled unmodified
(oracle/build_ref.py).  Inputs are seeded and stored next to the outputs, so the fixture is
self-contained; copy it to tests/golden/ref_cuda_ops.npz.  The CPU oracle (oracle/pn2_oracle.c)
and this repository's kernels are both checked against it.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.build_ref import load_ref_ext  # noqa: E402
from situation3d_b200.synthetic import make_scene  # noqa: E402


def cloud(g, b, n, scale=1.0, dup=0, zeros=0):
    x = torch.randn(b, n, 3, generator=g) * scale
    for bi in range(b):
        if dup:
            src = torch.randint(0, n, (dup,), generator=g)
            dst = torch.randint(0, n, (dup,), generator=g)
            x[bi, dst] = x[bi, src]
        if zeros:
            x[bi, torch.randint(0, n, (zeros,), generator=g)] = 0.0
    return x


def main():
    ext = load_ref_ext()
    assert ext is not None, "oracle/_ref/pn2_ref_ext.so missing (run oracle/build_ref.py where /root/reference exists)"
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(20261017)
    out = {}

    # furthest point sampling: block sizes 1..512, ties (duplicates), skipped near-origin points, m > n
    fps_cases = [("n1", 1, 1, 3), ("n7", 2, 7, 5), ("n300", 2, 300, 64), ("n512", 2, 512, 128),
                 ("n1000", 3, 1000, 256), ("n5000", 2, 5000, 512), ("m_gt_n", 1, 20, 40)]
    for name, b, n, m in fps_cases:
        x = cloud(g, b, n, dup=n // 10, zeros=min(3, n // 4))
        out["fps_%s_xyz" % name] = x.numpy()
        out["fps_%s_idx" % name] = ext.furthest_point_sampling(x.to(dev), m).cpu().numpy()
    x = torch.zeros(1, 64, 3)                                   # every point skipped -> all zeros
    x[0, :, 0] = torch.linspace(0, 0.02, 64)
    out["fps_allskip_xyz"] = x.numpy()
    out["fps_allskip_idx"] = ext.furthest_point_sampling(x.to(dev), 8).cpu().numpy()
    x = torch.randint(-3, 4, (2, 700, 3), generator=g).float() * 0.25   # lattice: massive exact ties
    out["fps_lattice_xyz"] = x.numpy()
    out["fps_lattice_idx"] = ext.furthest_point_sampling(x.to(dev), 200).cpu().numpy()
    scene = torch.from_numpy(make_scene(0, 40000, 0))[None, :, :3].contiguous()
    out["fps_scene0_idx"] = ext.furthest_point_sampling(scene.to(dev), 2048).cpu().numpy()
    out["fps_scene0_sum"] = np.array([float(scene.double().sum())])

    # ball query: empty balls, more hits than nsample, duplicates
    for name, b, n, m, r, ns in [("small", 2, 1000, 64, 0.5, 16), ("dense", 1, 2000, 32, 1.5, 8),
                                 ("sparse", 2, 300, 40, 0.05, 4)]:
        x = cloud(g, b, n, dup=n // 10)
        c = torch.cat([x[:, : m // 2], cloud(g, b, m - m // 2, scale=3.0)], dim=1).contiguous()
        out["bq_%s_xyz" % name], out["bq_%s_new" % name] = x.numpy(), c.numpy()
        out["bq_%s_args" % name] = np.array([r, ns], dtype=np.float64)
        out["bq_%s_idx" % name] = ext.ball_query(c.to(dev), x.to(dev), r, ns).cpu().numpy()
    sxyz = scene.to(dev)
    sidx = torch.from_numpy(out["fps_scene0_idx"]).to(dev)
    centres = ext.gather_points(sxyz.transpose(1, 2).contiguous(), sidx).transpose(1, 2).contiguous()
    out["bq_scene0_idx"] = ext.ball_query(centres, sxyz, 0.2, 64).cpu().numpy().astype(np.uint16)

    # three_nn: generic, ties, fewer than three known points
    for name, b, n, m in [("gen", 2, 100, 50), ("two", 1, 10, 2), ("one", 1, 5, 1)]:
        u, k = cloud(g, b, n), cloud(g, b, m, dup=m // 5)
        d2, idx = ext.three_nn(u.to(dev), k.to(dev))
        out["nn_%s_unknown" % name], out["nn_%s_known" % name] = u.numpy(), k.numpy()
        out["nn_%s_dist2" % name], out["nn_%s_idx" % name] = d2.cpu().numpy(), idx.cpu().numpy()

    # gather / group / interpolate and their gradients
    b, c, n, m, ns = 2, 5, 60, 17, 6
    f = torch.randn(b, c, n, generator=g)
    gi = torch.randint(0, n, (b, m), generator=g).int()
    out["ga_feat"], out["ga_idx"] = f.numpy(), gi.numpy()
    out["ga_out"] = ext.gather_points(f.to(dev), gi.to(dev)).cpu().numpy()
    go = torch.randn(b, c, m, generator=g)
    out["ga_gout"] = go.numpy()
    out["ga_grad"] = ext.gather_points_grad(go.to(dev), gi.to(dev), n).cpu().numpy()
    qi = torch.randint(0, n, (b, m, ns), generator=g).int()
    out["gr_idx"] = qi.numpy()
    out["gr_out"] = ext.group_points(f.to(dev), qi.to(dev)).cpu().numpy()
    go = torch.randn(b, c, m, ns, generator=g)
    out["gr_gout"] = go.numpy()
    out["gr_grad"] = ext.group_points_grad(go.to(dev), qi.to(dev), n).cpu().numpy()
    ti = torch.randint(0, n, (b, m, 3), generator=g).int()
    tw = torch.rand(b, m, 3, generator=g)
    out["ti_idx"], out["ti_w"] = ti.numpy(), tw.numpy()
    out["ti_out"] = ext.three_interpolate(f.to(dev), ti.to(dev), tw.to(dev)).cpu().numpy()
    go = torch.randn(b, c, m, generator=g)
    out["ti_gout"] = go.numpy()
    out["ti_grad"] = ext.three_interpolate_grad(go.to(dev), ti.to(dev), tw.to(dev), n).cpu().numpy()
    # the reference's own unit test fixture (pointnet2_test.py:18-30)
    feats = torch.randn(1, 2, 4, generator=g)
    idx = torch.tensor([[[0, 1, 2], [1, 2, 3]]], dtype=torch.int32)
    w = torch.tensor([[[1., 1., 1.], [2., 2., 2.]]])
    out["ut_feat"] = feats.numpy()
    out["ut_out"] = ext.three_interpolate(feats.to(dev), idx.to(dev), w.to(dev)).cpu().numpy()

    dst = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(dst, exist_ok=True)
    np.savez_compressed(os.path.join(dst, "ref_cuda_ops.npz"), **out)
    print("wrote", os.path.join(dst, "ref_cuda_ops.npz"), "with", len(out), "arrays")


if __name__ == "__main__":
    main()
