"""GPU parity suite, part 1: the nine operators through the reference-facing ``_ext`` API (which
calls the C ABI) against the CPU oracle, the committed golden vectors, the reference's own CUDA
kernels (oracle/_ref, when present) and size-independent properties at BASELINE.json sizes.
Integer / index outputs are compared bit for bit."""
import numpy as np
import pytest
import torch

from oracle import pn2_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    from situation3d_b200.pointnet2 import _ext
    from situation3d_b200 import _lib
    assert _lib.lib.pn2_device_check() == 0, "not an sm_100 device"
    return _ext


@pytest.fixture(scope="module")
def ref_ext():
    from oracle.build_ref import load_ref_ext
    mod = load_ref_ext()
    if mod is None:
        pytest.skip("oracle/_ref/pn2_ref_ext.so not built")
    return mod


def cloud(seed, b, n, dup=0, zeros=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, n, 3, generator=g) * scale
    for bi in range(b):
        if dup:
            x[bi, torch.randint(0, n, (dup,), generator=g)] = x[bi, torch.randint(0, n, (dup,), generator=g)]
        if zeros:
            x[bi, torch.randint(0, n, (zeros,), generator=g)] = 0.0
    return x


def scene_xyz(seeds, n=40000):
    from situation3d_b200.synthetic import make_scene
    return torch.from_numpy(np.stack([make_scene(s, n, 0)[:, :3] for s in seeds])).contiguous()


# ---- furthest point sampling --------------------------------------------------------------------

@pytest.mark.parametrize("b,n,m", [(1, 1, 1), (2, 7, 5), (3, 31, 31), (2, 300, 64), (2, 512, 128), (2, 513, 100),
                                   (3, 1000, 256), (2, 1024, 512), (2, 2048, 1024), (1, 4097, 300),
                                   (2, 5000, 512), (1, 20, 40), (2, 9000, 700), (1, 20000, 512)])
def test_fps_vs_oracle(ext, b, n, m):
    x = cloud(100 + n, b, n, dup=n // 10, zeros=min(3, n // 4))
    got = ext.furthest_point_sampling(x.cuda(), m).cpu()
    want = orc.furthest_point_sampling(x, m)
    assert torch.equal(got, want)


def test_fps_ties_and_skips(ext):
    g = torch.Generator().manual_seed(7)
    lattice = torch.randint(-3, 4, (2, 700, 3), generator=g).float() * 0.25     # massive exact ties
    assert torch.equal(ext.furthest_point_sampling(lattice.cuda(), 200).cpu(), orc.furthest_point_sampling(lattice, 200))
    skip = torch.zeros(1, 64, 3)
    skip[0, :, 0] = torch.linspace(0, 0.02, 64)                                  # |p|^2 <= 1e-3 everywhere
    assert torch.equal(ext.furthest_point_sampling(skip.cuda(), 8).cpu(), torch.zeros(1, 8, dtype=torch.int32))
    same = torch.ones(1, 100, 3)                                                 # all points identical
    assert torch.equal(ext.furthest_point_sampling(same.cuda(), 10).cpu(), orc.furthest_point_sampling(same, 10))


@pytest.mark.parametrize("cluster,threads", [(1, 512), (2, 512), (4, 256), (8, 512), (16, 256), (16, 512)])
def test_fps_every_cluster_shape(ext, monkeypatch, cluster, threads):
    monkeypatch.setenv("PN2_FPS_CLUSTER", str(cluster))
    monkeypatch.setenv("PN2_FPS_THREADS", str(threads))
    monkeypatch.setenv("PN2_FPS_BUCKET_MIN", "1000000000")      # the register-resident cluster kernels, not the bucketed one
    n = 6000 if cluster == 1 else 12000
    x = cloud(55, 2, n, dup=500, zeros=4)
    assert torch.equal(ext.furthest_point_sampling(x.cuda(), 300).cpu(), orc.furthest_point_sampling(x, 300))


@pytest.mark.parametrize("b,n,m", [(1, 1, 1), (2, 7, 5), (2, 33, 33), (2, 300, 64), (3, 1000, 256), (2, 2048, 1024),
                                   (1, 4097, 300), (1, 20, 40), (2, 9000, 700), (1, 20000, 512), (1, 70000, 300),
                                   (1, 140000, 200)])
def test_fps_bucketed_kernel_vs_oracle(ext, monkeypatch, b, n, m):
    """csrc/fps_bucket.cu (one CTA per scene, exact bounding-box pruning) forced for every size."""
    monkeypatch.setenv("PN2_FPS_BUCKET_MIN", "1")
    x = cloud(300 + n, b, n, dup=n // 10, zeros=min(3, n // 4))
    got = ext.furthest_point_sampling(x.cuda(), m).cpu()
    assert torch.equal(got, orc.furthest_point_sampling(x, m))


def test_fps_bucketed_kernel_edge_cases(ext, monkeypatch):
    monkeypatch.setenv("PN2_FPS_BUCKET_MIN", "1")
    g = torch.Generator().manual_seed(7)
    lattice = torch.randint(-3, 4, (2, 700, 3), generator=g).float() * 0.25     # massive exact ties
    assert torch.equal(ext.furthest_point_sampling(lattice.cuda(), 200).cpu(), orc.furthest_point_sampling(lattice, 200))
    big = torch.randint(-6, 7, (1, 12000, 3), generator=g).float() * 0.125      # ties across buckets and warps
    assert torch.equal(ext.furthest_point_sampling(big.cuda(), 400).cpu(), orc.furthest_point_sampling(big, 400))
    skip = torch.zeros(1, 64, 3)
    skip[0, :, 0] = torch.linspace(0, 0.02, 64)                                  # |p|^2 <= 1e-3 everywhere
    assert torch.equal(ext.furthest_point_sampling(skip.cuda(), 8).cpu(), torch.zeros(1, 8, dtype=torch.int32))
    same = torch.ones(1, 100, 3)                                                 # all points identical
    assert torch.equal(ext.furthest_point_sampling(same.cuda(), 10).cpu(), orc.furthest_point_sampling(same, 10))
    odd = cloud(11, 2, 3000, dup=50, zeros=3)                                    # non-finite coordinates
    odd[0, 5, 0] = float("nan")
    odd[0, 77] = float("inf")
    odd[1, 900, 2] = float("-inf")
    odd[1, 0, 1] = float("nan")                                                  # the unconditional first sample
    assert torch.equal(ext.furthest_point_sampling(odd.cuda(), 300).cpu(), orc.furthest_point_sampling(odd, 300))
    flat = cloud(12, 1, 5000)
    flat[..., 2] = 0.5                                                           # zero extent along one axis
    assert torch.equal(ext.furthest_point_sampling(flat.cuda(), 256).cpu(), orc.furthest_point_sampling(flat, 256))


def test_fps_bucketed_equals_cluster_kernel_full_size(ext, monkeypatch):
    x = scene_xyz([6, 7]).cuda()
    monkeypatch.setenv("PN2_FPS_BUCKET_MIN", "1000000000")
    a = ext.furthest_point_sampling(x, 2048)
    monkeypatch.setenv("PN2_FPS_BUCKET_MIN", "1")
    assert torch.equal(ext.furthest_point_sampling(x, 2048), a)
    from situation3d_b200 import fused
    inds, new_xyz = fused.fps_with_xyz(x, 2048)
    assert torch.equal(inds, a)
    assert torch.equal(new_xyz, torch.gather(x, 1, a.long()[..., None].expand(-1, -1, 3)))


def test_fps_full_scene_vs_oracle(ext):
    x = scene_xyz([0, 1])                       # BASELINE.json size: 40k points, 2048 samples
    got = ext.furthest_point_sampling(x.cuda(), 2048).cpu()
    assert torch.equal(got, orc.furthest_point_sampling(x, 2048))
    # properties: first index 0, all distinct (the scene has > 2048 distinct valid points)
    assert (got[:, 0] == 0).all()
    assert all(len(set(r.tolist())) == 2048 for r in got)


def test_fps_xyz_fusion(ext):
    from situation3d_b200 import fused
    x = cloud(3, 2, 3000, dup=100).cuda()
    inds, new_xyz = fused.fps_with_xyz(x, 256)
    assert torch.equal(inds, ext.furthest_point_sampling(x, 256))
    assert torch.equal(new_xyz, torch.gather(x, 1, inds.long()[..., None].expand(-1, -1, 3)))


def test_fps_stress_sizes(ext):
    # BASELINE.json config 5: 100k-200k points, npoint 4096 (oracle on one scene keeps the test short)
    x = scene_xyz([5], n=200000)
    got = ext.furthest_point_sampling(x.cuda(), 4096).cpu()
    assert torch.equal(got[:, :600], orc.furthest_point_sampling(x, 600))      # prefix property of greedy FPS
    assert len(set(got[0].tolist())) == 4096


# ---- ball query -------------------------------------------------------------------------------

@pytest.mark.parametrize("b,n,m,r,ns", [(2, 1000, 64, 0.5, 16), (1, 2000, 33, 1.5, 8), (2, 300, 40, 0.05, 4),
                                        (1, 130, 1, 10.0, 64), (2, 5000, 100, 0.3, 200), (1, 7, 3, 1.0, 5),
                                        (2, 2048, 1024, 0.4, 32)])
def test_ball_query_vs_oracle(ext, b, n, m, r, ns):
    x = cloud(200 + n, b, n, dup=n // 10)
    c = torch.cat([x[:, : m // 2], cloud(300 + m, b, m - m // 2, scale=3.0)], dim=1).contiguous()
    got = ext.ball_query(c.cuda(), x.cuda(), r, ns).cpu()
    assert torch.equal(got, orc.ball_query(c, x, r, ns))


@pytest.mark.parametrize("b,n,m,r,ns,scale", [
    (2, 4096, 300, 0.3, 16, 1.0), (1, 5000, 64, 0.05, 8, 1.0), (2, 6000, 128, 2.5, 64, 1.0),
    (1, 20000, 257, 0.15, 32, 3.0), (1, 4500, 40, 50.0, 128, 1.0), (1, 70000, 100, 0.2, 64, 2.0)])
def test_ball_query_grid_path_vs_oracle(ext, b, n, m, r, ns, scale):
    """n >= 4096 takes the uniform-grid kernel; results must equal the ascending scan bit for bit."""
    x = cloud(700 + n, b, n, dup=n // 10, zeros=5, scale=scale)
    c = torch.cat([x[:, : m // 2], cloud(800 + m, b, m - m // 2, scale=3.0 * scale)], dim=1).contiguous()
    got = ext.ball_query(c.cuda(), x.cuda(), r, ns).cpu()
    assert torch.equal(got, orc.ball_query(c, x, r, ns))


def test_ball_query_grid_degenerate_geometry(ext):
    g = torch.Generator().manual_seed(77)
    # planar cloud (zero extent in z), a lattice with many points exactly on cell boundaries, NaN / inf points
    plane = torch.rand(1, 5000, 3, generator=g) * 4
    plane[..., 2] = 1.25
    lattice = (torch.randint(0, 30, (1, 6000, 3), generator=g).float() * 0.2)
    bad = torch.randn(1, 5000, 3, generator=g)
    bad[0, 10] = float("nan")
    bad[0, 11, 1] = float("inf")
    for x, r in ((plane, 0.2), (lattice, 0.2), (lattice, 0.4), (bad, 0.3)):
        c = torch.cat([x[:, :50], x[:, :50] + 0.05, torch.full((1, 3, 3), 100.0)], dim=1).contiguous()
        got = ext.ball_query(c.cuda(), x.cuda(), r, 16).cpu()
        assert torch.equal(got, orc.ball_query(c, x, r, 16))


def test_ball_query_grid_built_ahead_from_pitched_rows(ext):
    """The two halves of the grid path (build from point_clouds rows on one stream, query once the centres
    exist) give the ascending-scan result; small scenes report that they take the plain scan."""
    from situation3d_b200 import fused
    x = cloud(901, 2, 6000, dup=500, zeros=5)
    rows = torch.cat([x, torch.randn(2, 6000, 5)], dim=2).contiguous().cuda()      # pitch 8
    c = torch.cat([x[:, :100], cloud(902, 2, 60, scale=3.0)], dim=1).contiguous()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        ws = fused.ball_query_grid_build(rows, 6000, c.shape[1], 0.3, 16)
    assert ws is not None
    torch.cuda.current_stream().wait_stream(side)
    got = fused.ball_query_grid_query(ws, 6000, c.cuda(), 0.3, 16).cpu()
    assert torch.equal(got, orc.ball_query(c, x, 0.3, 16))
    assert fused.ball_query_grid_build(rows[:, :1000].contiguous(), 1000, 10, 0.3, 16) is None


def test_ball_query_full_scene(ext):
    x = scene_xyz([2])
    inds = orc.furthest_point_sampling(x, 2048)
    centres = torch.gather(x, 1, inds.long()[..., None].expand(-1, -1, 3)).contiguous()
    got = ext.ball_query(centres.cuda(), x.cuda(), 0.2, 64).cpu()
    assert torch.equal(got, orc.ball_query(centres, x, 0.2, 64))
    # properties: a centre is one of the points, so its ball is never empty and rows are ascending up to the padding
    d = (x[0, got[0].long()] - centres[0, :, None, :]).norm(dim=-1)
    assert (d < 0.2 + 1e-6).all()


# ---- three_nn -----------------------------------------------------------------------------------

@pytest.mark.parametrize("b,n,m", [(2, 100, 50), (1, 10, 2), (1, 5, 1), (2, 512, 256), (2, 1024, 512), (1, 700, 1300)])
def test_three_nn_vs_oracle(ext, b, n, m):
    u, k = cloud(400 + n, b, n), cloud(500 + m, b, m, dup=m // 5)
    d2, idx = ext.three_nn(u.cuda(), k.cuda())
    wd2, widx = orc.three_nn(u, k)
    assert torch.equal(idx.cpu(), widx)
    assert torch.equal(d2.cpu(), wd2)                      # bit-exact, +inf for m < 3 included


# ---- gather / group / interpolate (+ grads) -----------------------------------------------------

def test_copy_ops_vs_oracle(ext):
    g = torch.Generator().manual_seed(9)
    b, c, n, m, ns = 3, 37, 500, 129, 9
    f = torch.randn(b, c, n, generator=g)
    gi = torch.randint(0, n, (b, m), generator=g).int()
    qi = torch.randint(0, n, (b, m, ns), generator=g).int()
    ti = torch.randint(0, n, (b, m, 3), generator=g).int()
    tw = torch.rand(b, m, 3, generator=g)
    assert torch.equal(ext.gather_points(f.cuda(), gi.cuda()).cpu(), orc.gather_points(f, gi))
    assert torch.equal(ext.group_points(f.cuda(), qi.cuda()).cpu(), orc.group_points(f, qi))
    assert torch.equal(ext.three_interpolate(f.cuda(), ti.cuda(), tw.cuda()).cpu(), orc.three_interpolate(f, ti, tw))
    go = torch.randn(b, c, m, generator=g)
    torch.testing.assert_close(ext.gather_points_grad(go.cuda(), gi.cuda(), n).cpu(), orc.gather_points_grad(go, gi, n),
                               rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ext.three_interpolate_grad(go.cuda(), ti.cuda(), tw.cuda(), n).cpu(),
                               orc.three_interpolate_grad(go, ti, tw, n), rtol=1e-5, atol=1e-5)
    go = torch.randn(b, c, m, ns, generator=g)
    torch.testing.assert_close(ext.group_points_grad(go.cuda(), qi.cuda(), n).cpu(), orc.group_points_grad(go, qi, n),
                               rtol=1e-5, atol=1e-5)


def test_reference_unit_test_gradcheck():
    # lib/pointnet2/pointnet2_test.py:18-30, same fixture and tolerances
    from torch.autograd import gradcheck
    from situation3d_b200.pointnet2 import pointnet2_utils
    feats = torch.randn(1, 2, 4).float().cuda().requires_grad_(True)

    def interpolate_func(inputs):
        idx = torch.from_numpy(np.array([[[0, 1, 2], [1, 2, 3]]])).int().cuda()
        weight = torch.from_numpy(np.array([[[1, 1, 1], [2, 2, 2]]])).float().cuda()
        return pointnet2_utils.three_interpolate(inputs, idx, weight)

    assert gradcheck(interpolate_func, feats, atol=1e-1, rtol=1e-1)


def test_empty_inputs(ext):
    z = torch.zeros(2, 0, 3).cuda()
    assert ext.ball_query(z, cloud(1, 2, 50).cuda(), 0.5, 4).shape == (2, 0, 4)
    assert ext.three_nn(z, cloud(1, 2, 50).cuda())[1].shape == (2, 0, 3)
    assert ext.gather_points(torch.zeros(2, 4, 10).cuda(), torch.zeros(2, 0, dtype=torch.int32).cuda()).shape == (2, 4, 0)
    assert torch.equal(ext.ball_query(cloud(1, 1, 5).cuda(), torch.zeros(1, 0, 3).cuda(), 0.5, 4).cpu(),
                       torch.zeros(1, 5, 4, dtype=torch.int32))


# ---- golden vectors and the reference's own kernels ----------------------------------------------

def test_golden_vectors_from_reference_cuda(ext, ref_cuda_golden):
    g = ref_cuda_golden
    for name in ["n1", "n7", "n300", "n512", "n1000", "n5000", "m_gt_n", "allskip", "lattice"]:
        want = g["fps_%s_idx" % name]
        got = ext.furthest_point_sampling(torch.from_numpy(g["fps_%s_xyz" % name]).cuda(), want.shape[1]).cpu().numpy()
        assert np.array_equal(got, want), name
    for name in ["small", "dense", "sparse"]:
        r, ns = g["bq_%s_args" % name]
        got = ext.ball_query(torch.from_numpy(g["bq_%s_new" % name]).cuda(), torch.from_numpy(g["bq_%s_xyz" % name]).cuda(),
                             float(r), int(ns)).cpu().numpy()
        assert np.array_equal(got, g["bq_%s_idx" % name]), name
    for name in ["gen", "two", "one"]:
        d2, idx = ext.three_nn(torch.from_numpy(g["nn_%s_unknown" % name]).cuda(),
                               torch.from_numpy(g["nn_%s_known" % name]).cuda())
        assert np.array_equal(idx.cpu().numpy(), g["nn_%s_idx" % name]) and \
            np.array_equal(d2.cpu().numpy(), g["nn_%s_dist2" % name]), name


def test_against_reference_cuda_kernels_full_size(ext, ref_ext):
    """Same inputs through the reference's unmodified kernels (oracle/_ref) and ours, at full size."""
    x = scene_xyz([3, 4]).cuda()
    ref_inds = ref_ext.furthest_point_sampling(x, 2048)
    assert torch.equal(ext.furthest_point_sampling(x, 2048), ref_inds)
    centres = ref_ext.gather_points(x.transpose(1, 2).contiguous(), ref_inds).transpose(1, 2).contiguous()
    assert torch.equal(ext.ball_query(centres, x, 0.2, 64), ref_ext.ball_query(centres, x, 0.2, 64))
    c2 = centres[:, :1024].contiguous()
    assert torch.equal(ext.ball_query(c2, centres, 0.4, 32), ref_ext.ball_query(c2, centres, 0.4, 32))
    d2, idx = ext.three_nn(centres, c2)
    rd2, ridx = ref_ext.three_nn(centres, c2)
    assert torch.equal(idx, ridx) and torch.equal(d2, rd2)
    f = torch.randn(2, 16, 40000, device="cuda")
    bq = ref_ext.ball_query(centres, x, 0.2, 64)
    assert torch.equal(ext.group_points(f, bq), ref_ext.group_points(f, bq))
    assert torch.equal(ext.gather_points(f, ref_inds), ref_ext.gather_points(f, ref_inds))
    w = torch.rand(2, 2048, 3, device="cuda")
    f2 = torch.randn(2, 16, 1024, device="cuda")
    assert torch.equal(ext.three_interpolate(f2, idx, w), ref_ext.three_interpolate(f2, idx, w))
