"""CPU suite, part 1: the oracle itself.

  * against the golden vectors produced by the reference's CUDA kernels (ref_cuda_ops.npz);
  * against the golden vectors produced by the reference's Python modules (ref_modules.npz);
  * against first-principles numpy restatements on tiny inputs (pure-Python loops);
  * the reference's only unit test (pointnet2_test.py:18-30) as a known-answer test.
"""
import numpy as np
import pytest
import torch

from oracle import pn2_oracle as orc


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


# ---- pinned to the reference CUDA kernels ---------------------------------------------------

FPS_CASES = ["n1", "n7", "n300", "n512", "n1000", "n5000", "m_gt_n", "allskip", "lattice"]


@pytest.mark.parametrize("name", FPS_CASES)
def test_fps_matches_reference_cuda(ref_cuda_golden, name):
    g = ref_cuda_golden
    want = g["fps_%s_idx" % name]
    got = orc.furthest_point_sampling(T(g["fps_%s_xyz" % name]), want.shape[1]).numpy()
    assert np.array_equal(got, want)


def test_fps_scene_matches_reference_cuda(ref_cuda_golden):
    from situation3d_b200.synthetic import make_scene
    scene = T(make_scene(0, 40000, 0))[None, :, :3].contiguous()
    if float(scene.double().sum()) != float(ref_cuda_golden["fps_scene0_sum"][0]):
        pytest.skip("synthetic scene differs from the one the golden was made from (numpy version)")
    got = orc.furthest_point_sampling(scene, 2048).numpy()
    assert np.array_equal(got, ref_cuda_golden["fps_scene0_idx"])
    centres = orc.gather_points(scene.transpose(1, 2).contiguous(), T(got)).transpose(1, 2).contiguous()
    bq = orc.ball_query(centres, scene, 0.2, 64).numpy()
    assert np.array_equal(bq, ref_cuda_golden["bq_scene0_idx"].astype(np.int32))


@pytest.mark.parametrize("name", ["small", "dense", "sparse"])
def test_ball_query_matches_reference_cuda(ref_cuda_golden, name):
    g = ref_cuda_golden
    r, ns = g["bq_%s_args" % name]
    got = orc.ball_query(T(g["bq_%s_new" % name]), T(g["bq_%s_xyz" % name]), float(r), int(ns)).numpy()
    assert np.array_equal(got, g["bq_%s_idx" % name])


@pytest.mark.parametrize("name", ["gen", "two", "one"])
def test_three_nn_matches_reference_cuda(ref_cuda_golden, name):
    g = ref_cuda_golden
    d2, idx = orc.three_nn(T(g["nn_%s_unknown" % name]), T(g["nn_%s_known" % name]))
    assert np.array_equal(idx.numpy(), g["nn_%s_idx" % name])
    assert np.array_equal(d2.numpy(), g["nn_%s_dist2" % name])      # bit-exact, inf included


def test_copy_ops_match_reference_cuda(ref_cuda_golden):
    g = ref_cuda_golden
    f = T(g["ga_feat"])
    assert np.array_equal(orc.gather_points(f, T(g["ga_idx"])).numpy(), g["ga_out"])
    assert np.array_equal(orc.group_points(f, T(g["gr_idx"])).numpy(), g["gr_out"])
    assert np.array_equal(orc.three_interpolate(f, T(g["ti_idx"]), T(g["ti_w"])).numpy(), g["ti_out"])
    n = f.shape[2]
    # gradients: the reference sums with float atomics in arbitrary order -> tolerance, not bits
    np.testing.assert_allclose(orc.gather_points_grad(T(g["ga_gout"]), T(g["ga_idx"]), n).numpy(), g["ga_grad"],
                               rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(orc.group_points_grad(T(g["gr_gout"]), T(g["gr_idx"]), n).numpy(), g["gr_grad"],
                               rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(orc.three_interpolate_grad(T(g["ti_gout"]), T(g["ti_idx"]), T(g["ti_w"]), n).numpy(),
                               g["ti_grad"], rtol=1e-5, atol=1e-6)


# ---- the reference's own unit test as a known answer ---------------------------------------

def test_reference_unit_test_known_answer():
    # pointnet2_test.py:18-30: idx=[[0,1,2],[1,2,3]], weight=[[1,1,1],[2,2,2]] => f0+f1+f2 and 2(f1+f2+f3)
    torch.manual_seed(0)
    feats = torch.randn(1, 2, 4)
    idx = torch.tensor([[[0, 1, 2], [1, 2, 3]]], dtype=torch.int32)
    w = torch.tensor([[[1., 1., 1.], [2., 2., 2.]]])
    out = orc.three_interpolate(feats, idx, w)
    want = torch.stack([feats[0, :, 0] + feats[0, :, 1] + feats[0, :, 2],
                        2 * (feats[0, :, 1] + feats[0, :, 2] + feats[0, :, 3])], dim=1)[None]
    torch.testing.assert_close(out, want, rtol=1e-6, atol=1e-6)
    grad = orc.three_interpolate_grad(torch.ones(1, 2, 2), idx, w, 4)
    torch.testing.assert_close(grad, torch.tensor([[[1., 3., 3., 2.]] * 2]))


# ---- first-principles restatements on tiny inputs ------------------------------------------

def _d2(a, b):
    # float32 FMUL, FFMA, FFMA with the x term first, emulated in float64 -> float32 steps
    d = (a.astype(np.float32) - b.astype(np.float32)).astype(np.float32)
    acc = np.float32(d[0] * d[0])
    acc = np.float32(np.float64(d[1]) * np.float64(d[1]) + np.float64(acc))
    return np.float32(np.float64(d[2]) * np.float64(d[2]) + np.float64(acc))


def _d2_yxz(a, b):
    # three_nn's contraction in the reference build: FMUL on the y term, then x, then z
    d = (a.astype(np.float32) - b.astype(np.float32)).astype(np.float32)
    acc = np.float32(d[1] * d[1])
    acc = np.float32(np.float64(d[0]) * np.float64(d[0]) + np.float64(acc))
    return np.float32(np.float64(d[2]) * np.float64(d[2]) + np.float64(acc))


def test_ball_query_first_principles():
    rng = np.random.default_rng(1)
    xyz = rng.normal(size=(1, 50, 3)).astype(np.float32)
    new = xyz[:, :7].copy()
    r, ns = 0.9, 5
    got = orc.ball_query(T(new), T(xyz), r, ns).numpy()
    r2 = np.float32(r) * np.float32(r)
    for j in range(7):
        hits = [k for k in range(50) if _d2(new[0, j], xyz[0, k]) < r2][:ns]
        want = (hits + [hits[0]] * (ns - len(hits))) if hits else [0] * ns
        assert got[0, j].tolist() == want


def test_fps_first_principles_small():
    # n <= 32 with block size = 2^floor(log2 n): literal simulation of the strided scan + tree
    rng = np.random.default_rng(2)
    for n, m in [(5, 4), (8, 8), (13, 6)]:
        xyz = rng.normal(size=(n, 3)).astype(np.float32)
        xyz[2] = xyz[1]                                  # exact duplicate -> tie
        bs = 1 << int(np.floor(np.log2(n)))
        temp = np.full(n, 1e10, np.float32)
        old, want = 0, [0]
        for _ in range(1, m):
            best = np.full(bs, -1, np.float32)
            besti = np.zeros(bs, np.int64)
            for t in range(bs):
                for k in range(t, n, bs):
                    p = xyz[k]
                    mag = np.float32(np.float64(p[2]) * p[2] + np.float32(np.float64(p[1]) * p[1] + np.float32(p[0] * p[0])))
                    if np.float64(mag) <= 1e-3:
                        continue
                    d2 = min(_d2(p, xyz[old]), temp[k])
                    temp[k] = d2
                    if d2 > best[t]:
                        best[t], besti[t] = d2, k
            s = bs // 2
            while s >= 1:
                for t in range(s):
                    if best[t + s] > best[t]:
                        best[t], besti[t] = best[t + s], besti[t + s]
                s //= 2
            old = int(besti[0])
            want.append(old)
        got = orc.furthest_point_sampling(T(xyz[None]), m).numpy()[0].tolist()
        assert got == want


def test_three_nn_first_principles():
    rng = np.random.default_rng(3)
    u = rng.normal(size=(1, 6, 3)).astype(np.float32)
    k = rng.normal(size=(1, 9, 3)).astype(np.float32)
    d2, idx = orc.three_nn(T(u), T(k))
    for j in range(6):
        d = np.array([_d2_yxz(u[0, j], k[0, i]) for i in range(9)])
        order = np.argsort(d, kind="stable")[:3]
        assert idx[0, j].tolist() == order.tolist()
        assert np.array_equal(d2[0, j].numpy(), d[order])


def test_opt_n_threads():
    lib = orc.lib()
    for n, want in [(1, 1), (2, 2), (3, 2), (255, 128), (256, 256), (511, 256), (512, 512), (40000, 512)]:
        assert lib.pn2o_opt_n_threads(n) == want


# ---- pinned to the reference Python modules -------------------------------------------------

def _sd(g, prefix):
    return {k[len(prefix):]: T(v) for k, v in g.items() if k.startswith(prefix)}


@pytest.mark.parametrize("name,kw", [
    ("sa_a", dict(npoint=32, radius=0.6, nsample=16, normalize_xyz=True)),
    ("sa_b", dict(npoint=20, radius=0.9, nsample=8, normalize_xyz=False)),
    ("sa_c", dict(npoint=16, radius=0.7, nsample=80, normalize_xyz=True))])
def test_sa_wiring_matches_reference_python(ref_modules_golden, name, kw):
    g = ref_modules_golden
    new_xyz, feats, inds, _ = orc.sa_module_votes(T(g[name + "_xyz"]), T(g[name + "_feats"]), _sd(g, name + "_sd_"),
                                                  "", use_xyz=True, **kw)
    assert np.array_equal(inds.numpy(), g[name + "_inds"])
    assert np.array_equal(new_xyz.numpy(), g[name + "_new_xyz"])
    np.testing.assert_allclose(feats.numpy(), g[name + "_new_feats"], rtol=1e-5, atol=1e-5)


def test_fp_wiring_matches_reference_python(ref_modules_golden):
    g = ref_modules_golden
    out = orc.fp_module(T(g["fp_unknown"]), T(g["fp_known"]), T(g["fp_uf"]), T(g["fp_kf"]), _sd(g, "fp_sd_"), "")
    np.testing.assert_allclose(out.numpy(), g["fp_out"], rtol=1e-5, atol=1e-5)


def test_query_and_group_matches_reference_python(ref_modules_golden):
    g = ref_modules_golden
    nf, gx, _ = orc.query_and_group(T(g["qg_xyz"]), T(g["qg_new_xyz"]), T(g["qg_feats"]), 0.8, 12,
                                    use_xyz=True, normalize_xyz=True)
    np.testing.assert_allclose(nf.numpy(), g["qg_out"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(gx.numpy(), g["qg_gxyz"], rtol=1e-6, atol=1e-6)


def test_situation_helpers_match_reference_python(ref_modules_golden):
    g = ref_modules_golden
    np.testing.assert_allclose(orc.quaternions_to_rotation_matrices(T(g["sit_quats"])).numpy(), g["sit_quat_R"],
                               rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(orc.batch_rotation_vector_to_matrix(T(g["sit_rotvec"])).numpy(), g["sit_rotvec_R"],
                               rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(orc.batch_matrix_function(T(g["sit_vec"])).numpy(), g["sit_M"], rtol=1e-6, atol=1e-6)
    out, new_pos, prior = orc.reencode(T(g["sit_tokens"]), T(g["sit_pos"]), T(g["sit_vec"]),
                                       T(g["sit_pe_0.weight"]), T(g["sit_pe_0.bias"]), T(g["sit_pe_2.weight"]),
                                       T(g["sit_pe_2.bias"]))
    np.testing.assert_allclose(new_pos.numpy(), g["sit_pos_t"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(prior.numpy(), g["sit_prior"], rtol=1e-5, atol=1e-7)
    # the release embeds the UNtransformed positions (sqa_module.py:319); with an identity situation the
    # oracle's embedding path must reproduce that
    ident = torch.zeros(6, 7)
    ident[:, 6] = 1
    out_id, _, _ = orc.reencode(T(g["sit_tokens"]), T(g["sit_pos"]), ident, T(g["sit_pe_0.weight"]),
                                T(g["sit_pe_0.bias"]), T(g["sit_pe_2.weight"]), T(g["sit_pe_2.bias"]))
    np.testing.assert_allclose(out_id.numpy(), g["sit_tokens_pe"], rtol=1e-5, atol=1e-5)


def _token_case(g, name):
    n = int(g[name + "_nscenes"])
    coords = [torch.from_numpy(g["%s_coords%d" % (name, i)]) for i in range(n)]
    feats = [torch.from_numpy(g["%s_feats%d" % (name, i)]) for i in range(n)]
    return coords, feats


@pytest.mark.parametrize("name", ["a", "b"])
def test_column_tokens_oracle_vs_reference_loop(name):
    """The oracle's restatement of sqa_module.py:297-317 against the outputs of the reference's own loop (executed
    unmodified by tests/golden/make_ref_token_goldens.py), RNG stream included: bit for bit."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_tokens.npz"))
    coords, feats = _token_case(g, name)
    torch.manual_seed(1234)
    tok, pos, uniq, red, inv = orc.column_tokens(coords, feats, [16, 16, 16], 256, 0.02)
    assert torch.equal(tok, torch.from_numpy(g[name + "_scene_feat"]))
    assert torch.equal(pos, torch.from_numpy(g[name + "_scene_positions"]))
    # the quirk the CUDA path reproduces: scatter_reduce_('mean') onto zeros divides by count + 1
    for c, f, u, r, iv in zip(coords, feats, uniq, red, inv):
        cnt = torch.bincount(iv, minlength=u.shape[0]).float()
        want = torch.zeros_like(r).index_add_(0, iv, f) / (cnt + 1)[:, None]
        torch.testing.assert_close(r, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", ["rand", "sin"])
def test_voxel_pe_oracle_vs_reference_statements(name):
    """The oracle's restatement of blip2_t5.py:107-118 / blip2_opt.py:93-104 against the outputs of the reference's
    own statements (executed unmodified by tests/golden/make_ref_voxel_pe_goldens.py): bit for bit, float
    coordinates truncated toward zero and negative indices wrapped as torch does."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_voxel_pe.npz"))
    feat, pc, table = g[name + "_pc_feat"], g[name + "_pc"], g[name + "_table"]
    assert np.array_equal(orc.voxel_pe_all_pcs(pc, table, 1408), g[name + "_all_pcs"])
    assert np.array_equal(orc.voxel_pe(feat, pc, table, "add"), g[name + "_add"])
    cat = orc.voxel_pe(feat, pc, table, "cat")
    assert np.array_equal(cat[:, :feat.shape[1]], feat) and np.array_equal(cat[:, feat.shape[1]:], g[name + "_all_pcs"])
    assert (g[name + "_all_pcs"][..., 1407] == 0).all()                    # 1408 = 3 * 469 + 1: the last channel stays zero
    with pytest.raises(IndexError):
        orc.voxel_pe(feat, np.full_like(pc, table.shape[0]), table)


@pytest.mark.parametrize("layout", ["concat", "interleave"])
def test_voxel_pe_table(layout):
    """The product's table builder (torch fp32, the package's operation order) against the oracle's float64
    restatement of positional_encodings.PositionalEncoding1D -- tolerance only: the package is not in the reference
    tree and its version is not pinned (parity unpinned for the table; it is an input of the kernel)."""
    from situation3d_b200.voxel_pe import sinusoid_table
    t = sinusoid_table(256, 469, layout)
    assert t.shape == (256, 469) and t.dtype == torch.float32
    # fp32 angles up to 255 rad carry ~1.5e-5 absolute error into sin / cos
    np.testing.assert_allclose(t.numpy(), orc.voxel_pe_table(256, 469, layout), rtol=0, atol=5e-5)
    if layout == "concat":
        assert torch.all(t[0, :235] == 0) and torch.all(t[0, 235:] == 1)      # sin(0) block, then cos(0); column 469 is cut
    else:
        assert torch.all(t[0, 0::2] == 0) and torch.all(t[0, 1::2] == 1)


def _projection_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_projection.npz"))


def test_projection_oracle_vs_reference_class():
    """oracle compute_projection / frustum planes / project against the reference's own ProjectionHelper run on the CPU
    (tests/golden/make_ref_projection_goldens.py): index lists identical for every view (one view sees nothing ->
    None), corners and normals to fp32 rounding (the reference's bmm / cross leave the summation order open)."""
    g = _projection_golden()
    intrinsic, (dmin, dmax, acc), dims = torch.from_numpy(g["intrinsic"]), [float(v) for v in g["params"]], [int(v) for v in g["image_dims"]]
    np.testing.assert_allclose(orc.corner_points(intrinsic, dmin, dmax, dims).numpy(), g["corner_points"], rtol=0, atol=0)
    points = torch.from_numpy(g["points"])
    n = points.shape[0]
    nviews = g["poses"].shape[0]
    seen_none = False
    for v in range(nviews):
        c2w = torch.from_numpy(g["poses"][v])
        res = orc.compute_projection(points, torch.from_numpy(g["depths"][v]), c2w, torch.inverse(c2w), intrinsic, dmin, dmax, dims, acc)
        want3, want2 = g["indices_3d"][v].astype(np.int64), g["indices_2d"][v].astype(np.int64)
        if want3[0] == 0:
            assert res is None
            seen_none = True
            continue
        assert np.array_equal(res[0].numpy(), want3) and np.array_equal(res[1].numpy(), want2)
        k = int(want3[0])
        assert (np.diff(want3[1:1 + k]) > 0).all() and (want3[1 + k:] == 0).all()           # ascending, zero-padded
        corners, normals = orc.frustum_planes(c2w, intrinsic, dmin, dmax, dims)
        np.testing.assert_allclose(corners.numpy(), g["corners"][v], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(normals.numpy(), g["normals"][v], rtol=1e-4, atol=1e-4)
    assert seen_none
    for v in range(nviews):                                  # points_in_frustum with the reference's own corners / normals
        want = np.unpackbits(g["frustum_masks"][v])[:n].astype(bool)
        assert np.array_equal(orc.points_in_frustum(g["corners"][v], g["normals"][v], points, True).numpy(), want)
        assert orc.points_in_frustum(g["corners"][v], g["normals"][v], points) == int(want.sum())
    out = orc.project(g["label"][3], g["indices_3d"][3], g["indices_2d"][3], n)
    assert np.array_equal(out, g["project_view3"])
    for v in range(nviews):
        s = orc.project(g["label"][v], g["indices_3d"][v], g["indices_2d"][v], n).astype(np.float64).sum()
        assert abs(s - g["project_checksum"][v]) <= 1e-9 * max(1.0, abs(g["project_checksum"][v]))


# ---- pinned to the reference's own Python modules, executed (not goldens) ---------------------------------------

def test_backbone_wiring_equals_reference_modules_byte_code():
    """oracle/ref_modules.py: the reference's unmodified pointnet2_modules / pointnet2_utils / pytorch_utils (byte code
    compiled from /root/reference into oracle/_ref/) composed as Pointnet2Backbone over the oracle's operators give
    exactly what the oracle's restated wiring (orc.backbone) gives, from the same state_dict."""
    import types
    from oracle import ref_modules, workload
    ref_modules.build()
    if not ref_modules.available():
        pytest.skip("oracle/_ref/*.pyc not built (needs /root/reference once)")
    ext = types.ModuleType("pn2_oracle_ext")
    for name in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
                 "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"):
        setattr(ext, name, getattr(orc, name))
    net = ref_modules.make_backbone(ref_modules.load(ext), 9).eval()
    sd = workload.backbone_state_dict(9, seed=3)
    assert not net.load_state_dict(sd).missing_keys
    pc = T(workload.synthetic().make_batch(2, 3000, 9))
    with torch.no_grad():
        got, want = net(pc), orc.backbone(pc, sd)
    for k in ("sa1_inds", "sa2_inds", "sa3_inds", "sa4_inds", "fp2_inds", "fp2_xyz", "fp2_features", "sa2_features"):
        assert torch.equal(got[k], want[k]), k


def test_reference_arm_workload_does_not_import_the_product():
    """bench.py --impl reference builds its scenes and weights through oracle/workload.py only."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from oracle import workload; workload.synthetic().make_scene(0, 64, 3); "
            "workload.backbone_state_dict(); assert not any(m.startswith('situation3d_b200') for m in sys.modules), "
            "[m for m in sys.modules if m.startswith('situation3d')]") % __import__("os").path.dirname(
        __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    subprocess.check_call([sys.executable, "-c", code])
