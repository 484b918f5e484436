"""GPU parity suite, part 3: the tcgen05 building blocks in isolation (shared-memory operand
layouts, matrix / instruction descriptors, TMEM load) and the bf16 fused SA kernel against a bf16
emulation of the reference wiring (same quantisation points, fp32 accumulation)."""
import numpy as np
import pytest
import torch

from oracle import pn2_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k", [(64, 64), (128, 128), (256, 128), (64, 144), (128, 272), (16, 16), (48, 48),
                                 (128, 32), (256, 256), (128, 208)])
def test_umma_selftest_gemm(n, k):
    from situation3d_b200._lib import check, lib, ptr, stream_ptr
    g = torch.Generator().manual_seed(n * 1000 + k)
    a = torch.randn(128, k, generator=g).bfloat16().cuda()
    b = torch.randn(n, k, generator=g).bfloat16().cuda()
    d = torch.zeros(128, n, device="cuda")
    check(lib.pn2_selftest_umma(n, k, ptr(a), ptr(b), ptr(d), stream_ptr()), "selftest_umma")
    want = a.float() @ b.float().t()
    torch.testing.assert_close(d, want, rtol=1e-4, atol=1e-3)


def _bf16(x):
    return x.bfloat16().float()


def _sa_bf16_emulation(xyz, feats, folded, npoint, radius, nsample):
    """The reference wiring (oracle ops for the indices) with the tensor-core kernel's quantisation:
    bf16 features / weights / inter-layer activations, xyz and layer-1 bias as bf16 hi+lo pairs."""
    inds = orc.furthest_point_sampling(xyz, npoint)
    new_xyz = orc.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    idx = orc.ball_query(new_xyz, xyz, radius, nsample)
    gx = (orc.group_points(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)) * (1.0 / radius)
    gx = _bf16(gx) + _bf16(gx - _bf16(gx))
    gf = orc.group_points(_bf16(feats).contiguous(), idx)
    (w1, b1), (w2, b2), (w3, b3) = folded
    b1q = _bf16(b1) + _bf16(b1 - _bf16(b1))
    from situation3d_b200.fused import split_first_layer
    if split_first_layer(feats.shape[1], w1.shape[0]):
        # split first layer (csrc/lin_tc.cu): the feature half is computed per point and rounded to bf16
        pp = _bf16(torch.einsum("oc,bcn->bon", _bf16(w1[:, 3:]), _bf16(feats)))
        x = torch.einsum("oc,bcps->bops", _bf16(w1[:, :3]), gx) + orc.group_points(pp.contiguous(), idx)
    else:
        x = torch.einsum("oc,bcps->bops", _bf16(w1[:, :3]), gx) + torch.einsum("oc,bcps->bops", _bf16(w1[:, 3:]), gf)
    x = _bf16(torch.relu(x + b1q[None, :, None, None]))
    x = _bf16(torch.relu(torch.einsum("oc,bcps->bops", _bf16(w2), x) + b2[None, :, None, None]))
    x = torch.relu(torch.einsum("oc,bcps->bops", _bf16(w3), x) + b3[None, :, None, None])
    return new_xyz, x.max(dim=3)[0], inds


@pytest.mark.parametrize("npoint,radius,nsample,mlp,n", [
    (256, 0.45, 64, [129, 64, 64, 128], 3000),
    (128, 0.8, 32, [128, 128, 128, 256], 600),
    (64, 1.2, 16, [256, 128, 128, 256], 300),
    (8, 2.0, 128, [16, 32, 48, 128], 500),
])
def test_sa_tc_vs_bf16_emulation(npoint, radius, nsample, mlp, n):
    """Tight check of the kernel's arithmetic: against an emulation with identical rounding points the only
    difference left is fp32 summation order."""
    from situation3d_b200.fused import fold_shared_mlp
    from situation3d_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes
    from situation3d_b200.synthetic import randomize_bn_stats
    torch.manual_seed(1)
    mod = randomize_bn_stats(PointnetSAModuleVotes(npoint=npoint, radius=radius, nsample=nsample, mlp=list(mlp),
                                                   normalize_xyz=True, precision="bf16")).eval()
    g = torch.Generator().manual_seed(11)
    xyz, feats = torch.randn(2, n, 3, generator=g), torch.randn(2, mlp[0], n, generator=g).relu()
    want_xyz, want, want_inds = _sa_bf16_emulation(xyz, feats, fold_shared_mlp(mod.mlp_module), npoint, radius, nsample)
    mod = mod.cuda()
    with torch.no_grad():
        new_xyz, out, inds = mod(xyz.cuda(), feats.cuda())
    assert torch.equal(inds.cpu(), want_inds) and torch.equal(new_xyz.cpu(), want_xyz)
    # a bf16 rounding of an intermediate activation can flip on a last-bit fp32 difference: allow 1 bf16 ulp
    # of the layer-2 activations to propagate (~4e-3 relative to the output scale)
    scale = want.abs().max()
    assert float((out.cpu() - want).abs().max()) <= 6e-3 * float(scale)


def test_sa_tc_unsupported_shapes_use_fp32_kernel():
    from situation3d_b200._lib import lib
    assert lib.pn2_sa_tc_supported(129, 64, 64, 128, 2048, 64) == 1
    assert lib.pn2_sa_tc_supported(128, 128, 128, 256, 1024, 32) == 1
    assert lib.pn2_sa_tc_supported(256, 128, 128, 256, 256, 16) == 1
    assert lib.pn2_sa_tc_supported(129, 64, 64, 128, 2048, 48) == 0        # nsample not a power of two
    assert lib.pn2_sa_tc_supported(129, 64, 64, 100, 2048, 64) == 0        # c3 not a multiple of 128
    assert lib.pn2_sa_tc_supported(129, 64, 64, 128, 3, 64) == 0           # npoint*nsample % 128 != 0
    from situation3d_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes
    from situation3d_b200.synthetic import randomize_bn_stats
    torch.manual_seed(2)
    mod = randomize_bn_stats(PointnetSAModuleVotes(npoint=20, radius=0.9, nsample=8, mlp=[6, 24, 40],
                                                   precision="bf16")).eval()
    g = torch.Generator().manual_seed(3)
    xyz, feats = torch.randn(2, 300, 3, generator=g), torch.randn(2, 6, 300, generator=g)
    sd = {k: v.clone() for k, v in mod.state_dict().items()}
    _, want, _, _ = orc.sa_module_votes(xyz, feats, sd, "", 20, 0.9, 8)
    with torch.no_grad():
        _, out, _ = mod.cuda()(xyz.cuda(), feats.cuda())
    torch.testing.assert_close(out.cpu(), want, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("n,m,ck,cs,mlp", [(300, 100, 256, 256, [512, 256, 256]), (1024, 512, 256, 256, [512, 256, 256]),
                                            (129, 40, 64, 128, [192, 64, 48]), (70, 3, 128, 0, [128, 128, 16])])
def test_fp_tc_vs_bf16_emulation(n, m, ck, cs, mlp):
    """Tensor-core FP kernel against the reference wiring evaluated with the same rounding points:
    bf16 inputs / weights / interpolated values / layer-1 activations, fp32 accumulation."""
    from situation3d_b200 import fused
    from situation3d_b200.pointnet2.pointnet2_modules import PointnetFPModule
    from situation3d_b200.synthetic import randomize_bn_stats
    torch.manual_seed(5)
    mod = randomize_bn_stats(PointnetFPModule(mlp=list(mlp), precision="bf16")).eval()
    g = torch.Generator().manual_seed(17)
    unknown, known = torch.randn(2, n, 3, generator=g), torch.randn(2, m, 3, generator=g)
    kf = _bf16(torch.randn(2, ck, m, generator=g).relu())
    uf = _bf16(torch.randn(2, cs, n, generator=g).relu()) if cs else None
    (w1, b1), (w2, b2) = fused.fold_shared_mlp(mod.mlp)
    d2, idx = orc.three_nn(unknown, known)
    recip = 1.0 / (torch.sqrt(d2) + 1e-8)
    w = recip / recip.sum(dim=2, keepdim=True)
    interp = _bf16(orc.three_interpolate(kf.contiguous(), idx, w))
    x = torch.cat([interp, uf], dim=1) if cs else interp
    h = _bf16(torch.relu(torch.einsum("oc,bcn->bon", _bf16(w1), x) + b1[None, :, None]))
    want = torch.relu(torch.einsum("oc,bcn->bon", _bf16(w2), h) + b2[None, :, None])
    mod = mod.cuda()
    img = mod._image.get(mod.mlp, "bf16", kind="fp")
    assert not img.f32_only
    dd2, didx = fused.three_nn(unknown.cuda(), known.cuda())
    known_rows = kf.transpose(1, 2).contiguous().bfloat16().cuda()
    skip_rows = uf.transpose(1, 2).contiguous().bfloat16().cuda() if cs else None
    out, out_rows = fused.fp_forward_bf16(img, dd2, didx, known_rows, skip_rows)
    scale = float(want.abs().max())
    assert float((out.cpu() - want).abs().max()) <= 6e-3 * scale
    torch.testing.assert_close(out_rows.float().cpu().transpose(1, 2), out.cpu(), rtol=1e-2, atol=1e-2 * scale)


@pytest.mark.parametrize("rows,c,skip,ld,c1,bf16_in", [
    (1000, 129, 3, 132, 64, False),      # SA1: point_clouds rows read in place
    (300, 61, 0, 64, 128, False),
    (128 * 5, 256, 0, 256, 128, True),   # SA3 / SA4: bf16 rows of the previous layer
    (77, 129, 0, 136, 64, True),
])
def test_lin_tc_row_gemm(rows, c, skip, ld, c1, bf16_in):
    """P = X W^T of the split first layer against the same product in PyTorch (bf16 operands, fp32 sum)."""
    from situation3d_b200 import fused
    from situation3d_b200._lib import check, lib, ptr, stream_ptr
    g = torch.Generator().manual_seed(rows + c)
    w = torch.randn(c1, 3 + c, generator=g).cuda()
    x = torch.randn(rows, ld, generator=g)
    if bf16_in:
        x[:, c:] = 0                       # padding columns of a bf16 row table are zero
        xd = x.bfloat16().cuda()
        kin = ld
    else:
        xd = x.cuda()
        kin = (skip + c + 3) // 4 * 4
    assert lib.pn2_lin_tc_supported(kin, c1, int(bf16_in))
    image = torch.empty(lib.pn2_lin_tc_weight_image_bytes(kin, c1), dtype=torch.uint8, device="cuda")
    check(lib.pn2_lin_tc_pack_weights(kin, skip, c, c1, ptr(w), 3 + c, 3, ptr(image), stream_ptr()), "pack")
    got = fused.lin_rows(image, xd, ptr(xd), rows, kin, c1, bf16_in, ld).float().cpu()
    want = (_bf16(x[:, skip:skip + c]) @ _bf16(w[:, 3:].cpu()).t())
    err = (got - want).abs()
    assert float((err - 2 ** -8 * want.abs()).max()) <= 1e-3      # one bf16 rounding of the result
