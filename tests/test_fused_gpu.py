"""GPU parity suite, part 2: the SA / FP modules, the backbone and the situation re-encoding
through the reference-facing Python API, against the CPU oracle's restatement of the reference
wiring and the golden vectors produced by the reference's own Python modules.

Tolerances (BASELINE.json north_star): indices bit-exact; features rtol 1e-3 for fp32,
2e-2 for bf16 (relative to the tensor's scale, see ``assert_features``)."""
import numpy as np
import pytest
import torch

from oracle import pn2_oracle as orc

pytestmark = pytest.mark.gpu

from oracle.parity import check_features


def assert_features(got, want, precision):
    """Acceptance criteria: oracle/parity.py (fp32 elementwise rtol 1e-3; bf16 2e-2 of the tensor range
    and 1e-2 relative L2)."""
    ok, msg = check_features(got, want, precision)
    assert ok, msg


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def precisions():
    from situation3d_b200 import fused
    return sorted(fused.SA_FORWARD.keys())


# ---- golden vectors made by the reference's Python modules ----------------------------------------

@pytest.mark.parametrize("name,kw", [
    ("sa_a", dict(npoint=32, radius=0.6, nsample=16, mlp=[6, 16, 16, 32], normalize_xyz=True)),
    ("sa_b", dict(npoint=20, radius=0.9, nsample=8, mlp=[6, 24, 40], normalize_xyz=False)),
    ("sa_c", dict(npoint=16, radius=0.7, nsample=80, mlp=[6, 16, 32], normalize_xyz=True))])
@pytest.mark.parametrize("fused", [True, False])
def test_sa_module_vs_reference_python_golden(ref_modules_golden, name, kw, fused):
    from situation3d_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes
    g = ref_modules_golden
    mod = PointnetSAModuleVotes(fused=fused, precision="fp32", **{k: (list(v) if isinstance(v, list) else v)
                                                                 for k, v in kw.items()})
    mod.load_state_dict({k[len(name) + 4:]: T(v) for k, v in g.items() if k.startswith(name + "_sd_")})
    mod = mod.cuda().eval()
    with torch.no_grad():
        new_xyz, feats, inds = mod(T(g[name + "_xyz"]).cuda(), T(g[name + "_feats"]).cuda())
    assert np.array_equal(inds.cpu().numpy(), g[name + "_inds"])
    assert np.array_equal(new_xyz.cpu().numpy(), g[name + "_new_xyz"])
    assert_features(feats, T(g[name + "_new_feats"]), "fp32")


@pytest.mark.parametrize("fused", [True, False])
def test_fp_module_vs_reference_python_golden(ref_modules_golden, fused):
    from situation3d_b200.pointnet2.pointnet2_modules import PointnetFPModule
    g = ref_modules_golden
    mod = PointnetFPModule(mlp=[44, 48, 24], fused=fused, precision="fp32")
    mod.load_state_dict({k[len("fp_sd_"):]: T(v) for k, v in g.items() if k.startswith("fp_sd_")})
    mod = mod.cuda().eval()
    with torch.no_grad():
        out = mod(T(g["fp_unknown"]).cuda(), T(g["fp_known"]).cuda(), T(g["fp_uf"]).cuda(), T(g["fp_kf"]).cuda())
    assert_features(out, T(g["fp_out"]), "fp32")


def test_query_and_group_vs_reference_python_golden(ref_modules_golden):
    from situation3d_b200.pointnet2.pointnet2_utils import QueryAndGroup
    g = ref_modules_golden
    grouper = QueryAndGroup(0.8, 12, use_xyz=True, ret_grouped_xyz=True, normalize_xyz=True)
    nf, gx = grouper(T(g["qg_xyz"]).cuda(), T(g["qg_new_xyz"]).cuda(), T(g["qg_feats"]).cuda())
    torch.testing.assert_close(nf.cpu(), T(g["qg_out"]), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(gx.cpu(), T(g["qg_gxyz"]), rtol=1e-6, atol=1e-6)


# ---- modules against the oracle wiring on seeded inputs -------------------------------------------

def _module_case(seed, n, c):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(2, n, 3, generator=g), torch.randn(2, c, n, generator=g).relu()


@pytest.mark.parametrize("precision", precisions())
@pytest.mark.parametrize("npoint,radius,nsample,mlp,n", [
    (256, 0.45, 64, [129, 64, 64, 128], 3000),       # SA1 shape
    (128, 0.8, 32, [128, 128, 128, 256], 600),       # SA2 shape
    (64, 1.2, 16, [256, 128, 128, 256], 300),        # SA3/SA4 shape
])
def test_sa_module_vs_oracle(precision, npoint, radius, nsample, mlp, n):
    from situation3d_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes
    from situation3d_b200.synthetic import randomize_bn_stats
    torch.manual_seed(1)
    mod = randomize_bn_stats(PointnetSAModuleVotes(npoint=npoint, radius=radius, nsample=nsample, mlp=list(mlp),
                                                   normalize_xyz=True, precision=precision)).eval()
    xyz, feats = _module_case(11, n, mlp[0])
    sd = {k: v.clone() for k, v in mod.state_dict().items()}
    want_xyz, want_feats, want_inds, _ = orc.sa_module_votes(xyz, feats, sd, "", npoint, radius, nsample,
                                                             use_xyz=True, normalize_xyz=True)
    mod = mod.cuda()
    with torch.no_grad():
        new_xyz, out, inds = mod(xyz.cuda(), feats.cuda())
    assert torch.equal(inds.cpu(), want_inds)
    assert torch.equal(new_xyz.cpu(), want_xyz)
    assert_features(out, want_feats, precision)


@pytest.mark.parametrize("precision", precisions())
def test_fp_module_vs_oracle(precision):
    from situation3d_b200.pointnet2.pointnet2_modules import PointnetFPModule
    from situation3d_b200.synthetic import randomize_bn_stats
    torch.manual_seed(2)
    mod = randomize_bn_stats(PointnetFPModule(mlp=[512, 256, 256], precision=precision)).eval()
    g = torch.Generator().manual_seed(5)
    unknown, known = torch.randn(2, 300, 3, generator=g), torch.randn(2, 100, 3, generator=g)
    uf, kf = torch.randn(2, 256, 300, generator=g).relu(), torch.randn(2, 256, 100, generator=g).relu()
    want = orc.fp_module(unknown, known, uf, kf, {k: v.clone() for k, v in mod.state_dict().items()}, "")
    mod = mod.cuda()
    with torch.no_grad():
        out = mod(unknown.cuda(), known.cuda(), uf.cuda(), kf.cuda())
    assert_features(out, want, precision)


def test_training_mode_path_and_gradients():
    """Training mode runs operator by operator with autograd through the *_grad kernels.  The same
    module evaluated with plain torch indexing (no custom op, autograd by PyTorch) must give the same
    loss and the same gradients w.r.t. the input features and every parameter."""
    import copy
    import torch.nn.functional as F
    from situation3d_b200.pointnet2.pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes
    torch.manual_seed(3)
    sa = PointnetSAModuleVotes(npoint=32, radius=0.7, nsample=8, mlp=[5, 16, 32], normalize_xyz=True).cuda().train()
    fp = PointnetFPModule(mlp=[32 + 5, 16]).cuda().train()
    sa2, fp2 = copy.deepcopy(sa), copy.deepcopy(fp)
    xyz, feats = _module_case(21, 200, 5)
    xyz = xyz.cuda()
    feats_a = feats.cuda().requires_grad_(True)
    feats_b = feats.cuda().requires_grad_(True)

    new_xyz, out, inds = sa(xyz, feats_a)
    loss_a = fp(xyz, new_xyz, feats_a, out).square().mean()
    loss_a.backward()

    # plain-torch twin: same indices, gathers written with advanced indexing
    from situation3d_b200.pointnet2 import _ext
    B = xyz.shape[0]
    bidx = torch.arange(B, device="cuda")[:, None, None]
    ball = _ext.ball_query(new_xyz, xyz, 0.7, 8).long()
    g_xyz = (xyz[bidx, ball] - new_xyz[:, :, None, :]) / 0.7                      # (B, np, ns, 3)
    g_feat = feats_b.transpose(1, 2)[bidx, ball]                                   # (B, np, ns, C)
    grouped = torch.cat([g_xyz, g_feat], dim=-1).permute(0, 3, 1, 2)
    pooled = sa2.mlp_module(grouped).max(dim=3)[0]
    d2, nn_idx = _ext.three_nn(xyz, new_xyz)
    recip = 1.0 / (d2.sqrt() + 1e-8)
    w = recip / recip.sum(dim=2, keepdim=True)
    gathered = pooled.transpose(1, 2)[torch.arange(B, device="cuda")[:, None, None], nn_idx.long()]   # (B, n, 3, C)
    interp = (gathered * w[..., None]).sum(dim=2).transpose(1, 2)
    loss_b = fp2.mlp(torch.cat([interp, feats_b], dim=1).unsqueeze(-1)).squeeze(-1).square().mean()
    loss_b.backward()

    torch.testing.assert_close(loss_a, loss_b, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(feats_a.grad, feats_b.grad, rtol=1e-3, atol=1e-6)
    for (n1, p1), (_, p2) in zip(list(sa.named_parameters()) + list(fp.named_parameters()),
                                 list(sa2.named_parameters()) + list(fp2.named_parameters())):
        assert p1.grad is not None, n1
        torch.testing.assert_close(p1.grad, p2.grad, rtol=2e-3, atol=1e-6, msg=n1)


# ---- backbone ------------------------------------------------------------------------------------

def _backbone_pair(precision, n_points, npoints, seed=0, batch=2, feat=129):
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.synthetic import make_batch, randomize_bn_stats
    torch.manual_seed(0)
    net = randomize_bn_stats(Pointnet2Backbone(input_feature_dim=feat, precision=precision, npoints=npoints)).eval()
    pc = torch.from_numpy(make_batch(batch, n_points, feat, first_seed=seed))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    layers = tuple((name, npnt, r, ns) for (name, _, r, ns), npnt in zip(orc.BACKBONE_LAYERS, npoints))
    return net, pc, sd, layers


def _check_backbone(out, want, precision):
    for k in ("sa1", "sa2", "sa3", "sa4"):
        assert torch.equal(out[k + "_inds"].cpu(), want[k + "_inds"]), k
        assert torch.equal(out[k + "_xyz"].cpu(), want[k + "_xyz"]), k
        assert_features(out[k + "_features"], want[k + "_features"], precision)
    assert torch.equal(out["fp2_inds"].cpu(), want["fp2_inds"])
    assert torch.equal(out["fp2_xyz"].cpu(), want["fp2_xyz"])
    assert_features(out["fp2_features"], want["fp2_features"], precision)


@pytest.mark.parametrize("precision", precisions())
def test_backbone_small_vs_oracle(precision):
    net, pc, sd, layers = _backbone_pair(precision, 4000, (512, 256, 128, 64))
    want = orc.backbone(pc, sd, layers)
    with torch.no_grad():
        out = net.cuda()({"point_clouds": pc.cuda()})
    _check_backbone(out, want, precision)
    assert out["fp2_features"].shape == (2, 256, 256)


@pytest.mark.parametrize("precision", precisions())
def test_backbone_full_size_vs_oracle(precision):
    """BASELINE.json config: 40k points, 129 feature channels, SA 2048/1024/512/256."""
    net, pc, sd, layers = _backbone_pair(precision, 40000, (2048, 1024, 512, 256), seed=7, batch=1)
    want = orc.backbone(pc, sd, layers)
    with torch.no_grad():
        out = net.cuda()({"point_clouds": pc.cuda()})
    _check_backbone(out, want, precision)
    assert out["fp2_features"].shape == (1, 256, 1024)
    assert torch.equal(out["fp2_inds"], out["sa1_inds"][:, :1024])


def test_backbone_fused_equals_unfused_full_size():
    net, pc, _, _ = _backbone_pair("fp32", 40000, (2048, 1024, 512, 256), seed=9, batch=2)
    net = net.cuda()
    with torch.no_grad():
        a = net({"point_clouds": pc.cuda()})
        net.fused = False
        for m in (net.sa1, net.sa2, net.sa3, net.sa4, net.fp1, net.fp2):
            m.fused = False
        b = net({"point_clouds": pc.cuda()})
    for k in ("sa1_inds", "sa2_inds", "sa3_inds", "sa4_inds", "fp2_inds"):
        assert torch.equal(a[k], b[k])
    assert_features(a["fp2_features"], b["fp2_features"], "fp32")


# ---- situation re-encoding -----------------------------------------------------------------------

def test_situation_helpers_vs_reference_python_golden(ref_modules_golden):
    from situation3d_b200 import reencode as re
    g = ref_modules_golden
    torch.testing.assert_close(re.quaternions_to_rotation_matrices(T(g["sit_quats"]).cuda()).cpu(), T(g["sit_quat_R"]),
                               rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(re.batch_rotation_vector_to_matrix(T(g["sit_rotvec"]).cuda()).cpu(),
                               T(g["sit_rotvec_R"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(re.batch_matrix_function(T(g["sit_vec"]).cuda()).cpu(), T(g["sit_M"]),
                               rtol=1e-5, atol=1e-6)
    w = [T(g["sit_pe_" + k]).cuda() for k in ("0.weight", "0.bias", "2.weight", "2.bias")]
    out, new_pos, prior = re.reencode_tokens(T(g["sit_tokens"]).cuda(), T(g["sit_pos"]).cuda(), T(g["sit_vec"]).cuda(), *w)
    torch.testing.assert_close(new_pos.cpu(), T(g["sit_pos_t"]), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(prior.cpu(), T(g["sit_prior"]), rtol=1e-4, atol=1e-7)
    ident = torch.zeros(6, 7)
    ident[:, 6] = 1
    out_id, _, _ = re.reencode_tokens(T(g["sit_tokens"]).cuda(), T(g["sit_pos"]).cuda(), ident.cuda(), *w)
    torch.testing.assert_close(out_id.cpu(), T(g["sit_tokens_pe"]), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("agent", [False, True])
def test_reencode_vs_oracle_config3_shape(agent):
    """BASELINE.json config 3 shape: B=32 scenes x 256 tokens x 256 channels."""
    from situation3d_b200.reencode import SituationReencoder
    from situation3d_b200.synthetic import make_situations
    torch.manual_seed(4)
    mod = SituationReencoder(to_agent_frame=agent)
    g = torch.Generator().manual_seed(6)
    tokens = torch.randn(32, 256, 256, generator=g)
    pos = torch.randn(32, 256, 3, generator=g) * 2
    sit = torch.from_numpy(make_situations(32))
    pe = mod.pos_embed
    want, want_pos, want_prior = orc.reencode(tokens, pos, sit, pe[0].weight.detach(), pe[0].bias.detach(),
                                              pe[2].weight.detach(), pe[2].bias.detach(), to_agent_frame=agent)
    mod = mod.cuda()
    d = mod({"scene_feat": tokens.cuda(), "scene_positions": pos.cuda(), "auxiliary_task": sit.cuda()})
    torch.testing.assert_close(d["scene_feat"].cpu(), want, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(d["scene_positions_agent"].cpu(), want_pos, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(d["auxiliary_task_loc_gt"].cpu(), want_prior, rtol=1e-3, atol=1e-7)
    torch.testing.assert_close(d["auxiliary_task_loc_gt"].sum(1).cpu(), torch.ones(32), rtol=1e-4, atol=1e-4)
    assert torch.equal(d["att_feat_pre"].cpu(), tokens)


def test_reencode_roundtrip_property():
    """Forward transform followed by the agent-frame (inverse) transform is the identity for unit quaternions."""
    from situation3d_b200.reencode import reencode_tokens
    from situation3d_b200.synthetic import make_situations
    g = torch.Generator().manual_seed(8)
    pos = (torch.randn(8, 100, 3, generator=g) * 3).cuda()
    sit = torch.from_numpy(make_situations(8, seed=3)).cuda()
    tok = torch.zeros(8, 100, 16).cuda()
    w1, b1, w2, b2 = torch.zeros(4, 2).cuda(), torch.zeros(4).cuda(), torch.zeros(16, 4).cuda(), torch.zeros(16).cuda()
    _, fwd, _ = reencode_tokens(tok, pos, sit, w1, b1, w2, b2, want_prior=False)
    _, back, _ = reencode_tokens(tok, fwd, sit, w1, b1, w2, b2, to_agent_frame=True, want_prior=False)
    torch.testing.assert_close(back, pos, rtol=1e-4, atol=1e-4)


def test_situated_scene_encoder_config3():
    """BASELINE.json config 3: backbone + agent-frame transform + positional embedding of 256 visual tokens."""
    from situation3d_b200.scene_encoder import SituatedSceneEncoder
    from situation3d_b200.synthetic import make_batch, make_situations, randomize_bn_stats
    torch.manual_seed(0)
    enc = randomize_bn_stats(SituatedSceneEncoder(129, 256, precision="fp32")).eval()
    pc = torch.from_numpy(make_batch(2, 40000, 129, first_seed=30))
    sit = torch.from_numpy(make_situations(2, seed=1))
    sd = {k[len("backbone_net."):]: v.clone() for k, v in enc.state_dict().items() if k.startswith("backbone_net.")}
    want = orc.backbone(pc, sd)
    tokens = want["fp2_features"][:, :, :256].transpose(1, 2).contiguous()
    pos = want["fp2_xyz"][:, :256].contiguous()
    pe = enc.reencoder.pos_embed
    want_tok, want_pos, want_prior = orc.reencode(tokens, pos, sit, pe[0].weight.detach(), pe[0].bias.detach(),
                                                  pe[2].weight.detach(), pe[2].bias.detach())
    enc = enc.cuda()
    with torch.no_grad():
        d = enc({"point_clouds": pc.cuda(), "auxiliary_task": sit.cuda()})
    assert torch.equal(d["scene_token_inds"].cpu(), want["sa1_inds"][:, :256])
    assert torch.equal(d["scene_positions"].cpu(), pos)
    assert_features(d["scene_feat"], want_tok, "fp32")
    torch.testing.assert_close(d["scene_positions_agent"].cpu(), want_pos, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(d["auxiliary_task_loc_gt"].cpu(), want_prior, rtol=1e-3, atol=1e-7)
    # VoteNet names of the seeds (lib/loss_helper.py:47-59) and, with hidden_size, SIG3D.scene_feat_linear on the tokens
    assert d["seed_xyz"] is d["fp2_xyz"] and d["seed_inds"] is d["fp2_inds"] and d["seed_features"] is d["fp2_features"]
    enc2 = SituatedSceneEncoder(129, 256, precision="bf16", hidden_size=768).eval().cuda()
    with torch.no_grad():
        d2 = enc2({"point_clouds": pc.cuda(), "auxiliary_task": sit.cuda()})
        lin = enc2.scene_feat_linear[0]
        want_h = torch.nn.functional.gelu(torch.nn.functional.linear(d2["scene_feat"], lin.weight, lin.bias))
    assert d2["scene_feat_hidden"].shape == (2, 256, 768)
    assert float((d2["scene_feat_hidden"] - want_h).abs().max()) <= 2e-2 * float(want_h.abs().max())


def test_backbone_stress_config5_shape():
    """BASELINE.json config 5 shape: 100k-point scene, SA1 npoint 4096 / nsample 64 (FPS on 16 CTAs with
    shared-memory-resident coordinates, grid ball query, tensor-core SA) against the oracle."""
    npoints = (4096, 2048, 1024, 512)
    net, pc, sd, layers = _backbone_pair("bf16", 100000, npoints, seed=40, batch=1)
    want = orc.backbone(pc, sd, layers)
    with torch.no_grad():
        out = net.cuda()({"point_clouds": pc.cuda()})
    _check_backbone(out, want, "bf16")
    assert out["fp2_features"].shape == (1, 256, 2048)


def test_training_step_config4_single_gpu():
    """BASELINE.json config 4 on one GPU (the all-reduce is a no-op at world size 1; the N > 1 exchange is
    covered by tests/test_dist_cpu.py over gloo): loss is finite, decreases, every parameter moves."""
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.synthetic import make_batch
    from situation3d_b200.train_step import BackboneTrainer
    torch.manual_seed(0)
    net = Pointnet2Backbone(input_feature_dim=9, npoints=(256, 128, 64, 32)).cuda()
    before = [p.detach().clone() for p in net.parameters()]
    tr = BackboneTrainer(net, lr=1e-2)
    pc = torch.from_numpy(make_batch(2, 3000, 9, first_seed=50)).cuda()
    losses = [float(tr.step(pc)) for _ in range(4)]
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0]
    assert all(not torch.equal(a, b) for a, b in zip(before, net.parameters()))
    for p_, v in zip(net.parameters(), tr.bucket.views):
        assert p_.grad.data_ptr() == v.data_ptr()


def test_launch_count_and_sm_partition():
    """pn2_launch_count counts this library's kernels (21 per bf16 backbone step of the bench shape family), and
    running the sampling chains / the rest on two SM partitions (CUDA green contexts) changes nothing in the results."""
    from situation3d_b200._lib import lib
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.streams import SmPartition
    from situation3d_b200.synthetic import make_batch, randomize_bn_stats
    torch.manual_seed(0)
    net = randomize_bn_stats(Pointnet2Backbone(input_feature_dim=129, precision="bf16")).eval().cuda()
    pc = torch.from_numpy(make_batch(2, 40000, 129, first_seed=7)).cuda()
    with torch.no_grad():
        want = net({"point_clouds": pc})
        torch.cuda.synchronize()
        n0 = lib.pn2_launch_count()
        net({"point_clouds": pc})
        torch.cuda.synchronize()
        assert lib.pn2_launch_count() - n0 == 21          # 23 before three_nn moved into the FP kernel (csrc/fp_tc2.cu)
        part = SmPartition(64)
        assert part.sms[0] >= 64 and part.sms[0] + part.sms[1] <= 148 and part.sms[1] > 0
        net.sm_partition = part
        net._side_streams.clear()
        lane = part.stream(part.MAIN)
        assert lib.pn2_stream_sm_count(lane.cuda_stream) == part.sms[1]
        lane.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(lane):
            got = net({"point_clouds": pc})
        lane.synchronize()
    for k in ("sa1_inds", "sa4_inds", "fp2_inds"):
        assert torch.equal(got[k], want[k])
    assert torch.equal(got["fp2_features"], want["fp2_features"])


def test_graphed_backbone_matches_eager():
    """CUDA-graph replay of the step (both streams captured) returns what the eager call returns, for new inputs too."""
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.graphs import GraphedBackbone
    from situation3d_b200.synthetic import make_batch, randomize_bn_stats
    torch.manual_seed(0)
    net = randomize_bn_stats(Pointnet2Backbone(input_feature_dim=129, precision="bf16")).eval().cuda()
    pcs = [torch.from_numpy(make_batch(2, 40000, 129, first_seed=s)).cuda() for s in (3, 5)]
    with torch.no_grad():
        want = [{k: v.clone() for k, v in net({"point_clouds": pc}).items() if k != "point_clouds"} for pc in pcs]
        step = GraphedBackbone(net, pcs[0])
        assert step.launches_per_replay == 21
        for pc, w in zip(pcs, want):
            out = step(pc)
            step.stream.synchronize()
            for k in ("sa1_inds", "sa2_inds", "sa3_inds", "sa4_inds", "fp2_inds", "fp2_xyz", "fp2_features", "sa1_features"):
                assert torch.equal(out[k], w[k]), k


@pytest.mark.parametrize("name", ["a", "b"])
def test_column_tokens_vs_reference_loop(name):
    """Token construction (csrc/tokens.cu) against the outputs of the reference's own loop (sqa_module.py:297-315,
    executed unmodified by tests/golden/make_ref_token_goldens.py) with the same RNG seed: the sampled columns and
    their positions are identical, the pooled features equal the reference's sum / (count + 1)."""
    import os
    from situation3d_b200 import tokens
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_tokens.npz"))
    n = int(g[name + "_nscenes"])
    coords = [torch.from_numpy(g["%s_coords%d" % (name, i)]) for i in range(n)]
    feats = [torch.from_numpy(g["%s_feats%d" % (name, i)]) for i in range(n)]
    torch.manual_seed(1234)
    want_tok, want_pos, uniq, red, inv = orc.column_tokens(coords, feats, [16, 16, 16], 256, 0.02)
    assert torch.equal(want_tok, torch.from_numpy(g[name + "_scene_feat"]))         # oracle == reference (CPU suite too)
    cc, ff = [c.cuda() for c in coords], [f.cuda() for f in feats]
    offsets, ncols, pc, pf, inverse = tokens.column_pool(cc, ff)
    off = offsets.cpu().tolist()
    for s in range(n):
        assert ncols[s] == uniq[s].shape[0]
        assert torch.equal(pc[off[s]:off[s] + ncols[s]].cpu(), uniq[s].to(torch.int32))      # torch.unique(dim=0) order
        assert torch.equal(inverse[off[s]:off[s + 1]].cpu().long(), inv[s])                  # return_inverse
        torch.testing.assert_close(pf[off[s]:off[s] + ncols[s]].cpu(), red[s], rtol=1e-6, atol=1e-7)
    torch.manual_seed(1234)
    tok, pos = tokens.column_tokens(cc, ff, [16, 16, 16], 256, 0.02)
    assert torch.equal(pos.cpu(), torch.from_numpy(g[name + "_scene_positions"]))
    torch.testing.assert_close(tok.cpu(), torch.from_numpy(g[name + "_scene_feat"]), rtol=1e-6, atol=1e-7)


def test_column_pool_limits():
    from situation3d_b200 import tokens
    c = torch.zeros(4, 3, dtype=torch.int32, device="cuda")
    c[0, 0] = 1 << 20
    with pytest.raises(RuntimeError):
        tokens.column_pool([c], [torch.zeros(4, 8, device="cuda")])
    one = torch.tensor([[3, -5, 7]], dtype=torch.int32, device="cuda")
    offsets, ncols, pc, pf, inv = tokens.column_pool([one], [torch.full((1, 8), 2.0, device="cuda")])
    assert ncols == [1] and pc[0].tolist() == [3, -5] and torch.allclose(pf[0], torch.full((8,), 1.0, device="cuda"))


@pytest.mark.parametrize("B,T,k,n", [(32, 256, 256, 768), (3, 50, 64, 128), (1, 1, 128, 256)])
def test_scene_feat_linear_vs_torch(B, T, k, n):
    """SIG3D.scene_feat_linear (Linear + exact GELU, sqa_module.py:180-183) as one tcgen05 kernel against the PyTorch
    modules the reference runs: bf16 operands -> max|err| <= 2e-2 max|ref| (north_star's bf16 tolerance); state-dict keys
    are the reference's."""
    from situation3d_b200.heads import SceneFeatLinear
    torch.manual_seed(3)
    m = SceneFeatLinear(k, n).eval()
    assert list(m.state_dict().keys()) == ["0.weight", "0.bias"]
    x = torch.randn(B, T, k)
    with torch.no_grad():
        want = torch.nn.functional.gelu(torch.nn.functional.linear(x, m[0].weight, m[0].bias))
        got = m.cuda()(x.cuda())
        # tight check against the same product with bf16-rounded operands (what the kernel computes)
        wq = torch.nn.functional.gelu(torch.nn.functional.linear(x.bfloat16().float(), m[0].weight.cpu().bfloat16().float(), m[0].bias.cpu()))
    assert got.shape == (B, T, n)
    assert float((got.cpu() - want).abs().max()) <= 2e-2 * float(want.abs().max())
    torch.testing.assert_close(got.cpu(), wq, rtol=1e-4, atol=1e-4)
    # training mode runs the reference's modules (autograd)
    m.train()
    y = m(x.cuda().requires_grad_())
    y.sum().backward()
    assert m[0].weight.grad is not None


def test_backbone_pipeline_from_host_buffers():
    """BackbonePipeline (the serving loop bench.py times as e2e): batches submitted from pinned host memory come back
    in pinned host memory, equal to the eager forward, with more submissions than lanes."""
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.graphs import BackbonePipeline
    from situation3d_b200.synthetic import make_batch, randomize_bn_stats
    torch.manual_seed(0)
    net = randomize_bn_stats(Pointnet2Backbone(input_feature_dim=129, precision="bf16")).eval().cuda()
    host = [torch.from_numpy(make_batch(2, 40000, 129, first_seed=s)).pin_memory() for s in (11, 12, 13, 14, 15)]
    with torch.no_grad():
        want = [{k: net({"point_clouds": h.cuda()})[k].cpu() for k in ("fp2_features", "fp2_xyz", "fp2_inds")} for h in host]
        pipe = BackbonePipeline(net, host[0].cuda(), lanes=2)
        assert pipe.h2d_bytes == host[0].numel() * 4
        tickets = []
        for i, h in enumerate(host):
            if i >= 2:                                  # the lane is about to be reused: consume its previous result first
                t_prev, j = tickets[i - 2]
                got = pipe.result(t_prev)
                for k in want[j]:
                    assert torch.equal(got[k], want[j][k]), (j, k)
            tickets.append((pipe.submit(h), i))
        for t, j in tickets[-2:]:
            got = pipe.result(t)
            for k in want[j]:
                assert torch.equal(got[k], want[j][k]), (j, k)
        pipe.drain()


def test_row_layout_training_matches_reference_wiring():
    """Training mode on channel-last rows (train_rows.py) against the operator-by-operator wiring of the reference
    (QueryAndGroup -> NCHW SharedMLP -> max_pool2d, autograd through the *_grad kernels): same loss, same gradients
    for every parameter, same BatchNorm running statistics, same indices."""
    import copy
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.synthetic import make_batch
    torch.manual_seed(0)
    a = Pointnet2Backbone(input_feature_dim=9, npoints=(256, 128, 64, 32)).cuda().train()
    b = copy.deepcopy(a)
    b.train_layout = "reference"
    pc = torch.from_numpy(make_batch(2, 3000, 9)).cuda()
    oa, ob = a({"point_clouds": pc}), b({"point_clouds": pc})
    for k in ("sa1_inds", "sa2_inds", "sa3_inds", "sa4_inds", "fp2_inds"):
        assert torch.equal(oa[k], ob[k]), k
    torch.testing.assert_close(oa["fp2_features"], ob["fp2_features"], rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(oa["sa2_features"], ob["sa2_features"], rtol=1e-3, atol=1e-4)
    la, lb = oa["fp2_features"].square().mean(), ob["fp2_features"].square().mean()
    la.backward()
    lb.backward()
    torch.testing.assert_close(la, lb, rtol=1e-4, atol=1e-6)
    for (n1, p1), (_, p2) in zip(a.named_parameters(), b.named_parameters()):
        assert p1.grad is not None, n1
        scale = float(p2.grad.abs().max()) + 1e-12
        assert float((p1.grad - p2.grad).abs().max()) <= 5e-3 * scale, n1
    for (n1, t1), (_, t2) in zip(a.named_buffers(), b.named_buffers()):
        if t1.dtype.is_floating_point:
            torch.testing.assert_close(t1, t2, rtol=1e-3, atol=1e-5, msg=n1)
        else:
            assert torch.equal(t1, t2), n1


@pytest.mark.parametrize("groups,ns,c", [(4096, 0, 64), (300, 16, 128), (257, 32, 256), (1000, 64, 64), (50, 1, 128)])
def test_rows_bn_relu_pool_kernels_match_torch_autograd(groups, ns, c):
    """csrc/train_rows.cu (training-mode BatchNorm + ReLU [+ max over nsample] on rows, forward and backward) against
    the reference's chain BatchNorm2d -> ReLU -> max_pool2d under torch autograd, duplicated rows (ball-query padding:
    exact ties in the max) included."""
    import torch.nn.functional as F
    from situation3d_b200.train_rows import _BnReluRows
    g = torch.Generator().manual_seed(groups + ns + c)
    rows = groups * max(ns, 1)
    x = (torch.randn(rows, c, generator=g) * 2.0 + torch.randn(c, generator=g)).cuda()
    if ns > 1:
        x.view(groups, ns, c)[::3, ns // 2:] = x.view(groups, ns, c)[::3, :1]          # padded copies of the first neighbour
    gamma, beta = (torch.rand(c, generator=g) + 0.5).cuda(), torch.randn(c, generator=g).cuda()
    xa, ga, ba = (t.clone().requires_grad_(True) for t in (x, gamma, beta))
    xb, gb, bb = (t.clone().requires_grad_(True) for t in (x, gamma, beta))
    rm0, rv0 = torch.randn(c, generator=g).cuda(), (torch.rand(c, generator=g) + 0.5).cuda()
    rm, rv, rma, rva = rm0.clone(), rv0.clone(), rm0.clone(), rv0.clone()
    out = _BnReluRows.apply(xa, ga, ba, 1e-5, ns, 0.1, rma, rva)
    y = F.relu(F.batch_norm(xb, rm, rv, gb, bb, True, 0.1, 1e-5))
    if ns:
        y = F.max_pool2d(y.view(1, groups, ns, c).permute(0, 3, 1, 2), kernel_size=[1, ns]).squeeze(3).squeeze(0).t()
    torch.testing.assert_close(out, y, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rma, rm, rtol=1e-5, atol=1e-6)           # running statistics, as nn.BatchNorm2d updates them
    torch.testing.assert_close(rva, rv, rtol=1e-5, atol=1e-6)
    dout = torch.randn(out.shape, generator=g).cuda()
    out.backward(dout)
    y.backward(dout)
    for got, want, name in ((xa.grad, xb.grad, "dx"), (ga.grad, gb.grad, "dgamma"), (ba.grad, bb.grad, "dbeta")):
        scale = float(want.abs().max()) + 1e-12
        assert float((got - want).abs().max()) <= 2e-5 * scale + 1e-6, name


def test_group_rows_kernel_matches_query_and_group_on_rows():
    """pn2_group_rows (one pass) against the torch formulation of QueryAndGroup on rows (index_select, -=, /=, cat),
    forward bit for bit, backward = row scatter-add of the feature columns; source rows taken as a column slice of a
    wider matrix (point_clouds[..., 3:]) and as a contiguous tensor."""
    from situation3d_b200 import fused
    from situation3d_b200.train_rows import _GroupRows
    g = torch.Generator().manual_seed(5)
    B, N, C, npoint, ns, radius = 2, 3000, 9, 128, 16, 0.3
    pc = torch.randn(B, N, 3 + C, generator=g).cuda()
    xyz = pc[..., :3].contiguous()
    inds, new_xyz = fused.fps_with_xyz(xyz, npoint)
    idx = fused.ball_query(xyz, new_xyz, radius, ns)
    flat = (idx.long() + (torch.arange(B, device="cuda") * N)[:, None, None]).reshape(-1)
    for rows in (pc[..., 3:], pc[..., 3:].contiguous()):
        for normalize in (True, False):
            ra = rows.detach().clone().requires_grad_(True) if rows.is_contiguous() else rows
            got = _GroupRows.apply(ra, xyz, new_xyz, idx, radius, normalize)
            gx = xyz.reshape(B * N, 3).index_select(0, flat).view(B, npoint, ns, 3) - new_xyz[:, :, None, :]
            if normalize:
                gx = gx / radius
            rb = rows.detach().clone().requires_grad_(True)
            want = torch.cat([gx.reshape(-1, 3), rb.reshape(B * N, C).index_select(0, flat)], dim=1)
            assert torch.equal(got, want)
            if ra.requires_grad:
                dout = torch.randn(want.shape, generator=g).cuda()
                got.backward(dout)
                want.backward(dout)
                torch.testing.assert_close(ra.grad, rb.grad, rtol=1e-5, atol=1e-5)


def test_sampling_modes_give_identical_results():
    """fused.sampling_mode("throughput") swaps the FPS kernel of 40 000-point scenes (bucketed one-CTA-per-scene kernel
    instead of the 8-CTA cluster kernel); every output of the backbone stays bit-identical, eagerly and in a captured step."""
    from situation3d_b200 import fused
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.graphs import GraphedBackbone, pick_sampling_mode
    from situation3d_b200.synthetic import make_batch, randomize_bn_stats
    from situation3d_b200._lib import lib
    assert lib.pn2_furthest_point_sampling_workspace_bytes_mode(2, 40000, 2048, 0) == 0
    assert lib.pn2_furthest_point_sampling_workspace_bytes_mode(2, 40000, 2048, 1) > 0
    assert pick_sampling_mode(8) == "latency" and pick_sampling_mode(160) == "throughput"
    torch.manual_seed(0)
    net = randomize_bn_stats(Pointnet2Backbone(input_feature_dim=9, precision="bf16")).eval().cuda()
    pc = torch.from_numpy(make_batch(2, 40000, 9)).cuda()
    with torch.no_grad():
        a = net({"point_clouds": pc})
        with fused.sampling_mode("throughput"):
            b = net({"point_clouds": pc})
        g = GraphedBackbone(net, pc, sampling_mode="throughput")
        c = g(pc)
        torch.cuda.synchronize()
    for k in ("sa1_inds", "sa2_inds", "sa3_inds", "sa4_inds", "fp2_inds", "sa1_xyz", "fp2_xyz", "fp2_features"):
        assert torch.equal(a[k], b[k]), k
        assert torch.equal(a[k], c[k]), k


def test_compact_input_is_bit_identical():
    """fp32 coordinates + bf16 feature rows (Pointnet2Backbone.pack_point_clouds: half the host-to-device bytes) give
    bit-identical results to the reference's fp32 point_clouds on the bf16 arm, eagerly and through the pipeline from
    pinned host buffers; the format is refused where it would change results (fp32 arm, training)."""
    from situation3d_b200.backbone_module import Pointnet2Backbone
    from situation3d_b200.graphs import BackbonePipeline
    from situation3d_b200.synthetic import make_batch, randomize_bn_stats
    torch.manual_seed(0)
    net = randomize_bn_stats(Pointnet2Backbone(input_feature_dim=129, precision="bf16")).eval().cuda()
    host = [torch.from_numpy(make_batch(2, 40000, 129, first_seed=s)) for s in (21, 22, 23)]
    keys = ("sa1_inds", "sa2_inds", "sa3_inds", "sa4_inds", "fp2_inds", "fp2_xyz", "fp2_features", "sa1_features",
            "sa2_features", "sa4_features")
    with torch.no_grad():
        want = [{k: net({"point_clouds": h.cuda()})[k].cpu() for k in keys} for h in host]
        packed = [net.pack_point_clouds(h) for h in host]                      # on the host, like a data loader would
        assert packed[0][1].dtype == torch.bfloat16 and packed[0][1].shape == (2, 40000, 136)
        for (xyz, rows), w in zip(packed, want):
            out = net({"xyz": xyz.cuda(), "feature_rows_bf16": rows.cuda()})
            for k in keys:
                assert torch.equal(out[k].cpu(), w[k]), k
        pinned = [(x.pin_memory(), r.pin_memory()) for x, r in packed]
        pipe = BackbonePipeline(net, (pinned[0][0].cuda(), pinned[0][1].cuda()), lanes=2)
        assert pipe.h2d_bytes == 2 * 40000 * (12 + 136 * 2)
        for j, pr in enumerate(pinned):
            got = pipe.result(pipe.submit(pr))
            for k in ("fp2_features", "fp2_xyz", "fp2_inds"):
                assert torch.equal(got[k], want[j][k]), (j, k)
        pipe.drain()
        f32 = randomize_bn_stats(Pointnet2Backbone(input_feature_dim=129, precision="fp32")).eval().cuda()
        with pytest.raises(RuntimeError):
            f32({"xyz": packed[0][0].cuda(), "feature_rows_bf16": packed[0][1].cuda()})


@pytest.mark.parametrize("name", ["rand", "sin"])
def test_voxel_pe_vs_reference_statements(name):
    """csrc/voxel_pe.cu against the outputs of the reference's own statements (blip2_t5.py:107-118 and
    blip2_opt.py:93-104, executed unmodified by tests/golden/make_ref_voxel_pe_goldens.py): bit for bit in both
    modes, for float coordinates (the reference's `.long()` done in the kernel) and for int64 / int32 ones."""
    import os
    from situation3d_b200.voxel_pe import voxel_pe
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_voxel_pe.npz"))
    feat, pc, table = (torch.from_numpy(g[name + k]).cuda() for k in ("_pc_feat", "_pc", "_table"))
    want_add, all_pcs = torch.from_numpy(g[name + "_add"]), torch.from_numpy(g[name + "_all_pcs"])
    for coords in (pc, pc.long(), pc.int()):
        assert torch.equal(voxel_pe(feat, coords, table, "add").cpu(), want_add)
        cat = voxel_pe(feat, coords, table, "cat").cpu()
        assert torch.equal(cat[:, :feat.shape[1]], feat.cpu()) and torch.equal(cat[:, feat.shape[1]:], all_pcs)
    # in place, and a wider coordinate row (only the first three columns are read)
    wide = torch.cat([pc, torch.full_like(pc[..., :1], 1e9)], dim=-1)
    buf = feat.clone()
    assert voxel_pe(buf, wide, table, "add", out=buf) is buf and torch.equal(buf.cpu(), want_add)


@pytest.mark.parametrize("B,P,C,S,R", [(2, 5000, 1408, 469, 256), (1, 37, 30, 7, 5), (2, 33, 32, 9, 11), (3, 1, 1407, 469, 256), (1, 300, 1409, 469, 16)])
def test_voxel_pe_vs_oracle(B, P, C, S, R):
    """Reference-size batch and ragged shapes (C not a multiple of 4 takes the scalar path) against the numpy oracle."""
    from situation3d_b200.voxel_pe import voxel_pe
    g = torch.Generator().manual_seed(B * 1000 + P)
    feat = torch.randn(B, P, C, generator=g)
    pc = torch.randint(-R, R, (B, P, 3), generator=g)
    table = torch.randn(R, S, generator=g)
    for mode in ("add", "cat"):
        got = voxel_pe(feat.cuda(), pc.cuda(), table.cuda(), mode).cpu().numpy()
        assert np.array_equal(got, orc.voxel_pe(feat.numpy(), pc.numpy(), table.numpy(), mode)), mode


def test_voxel_pe_module_and_errors():
    from situation3d_b200.voxel_pe import VoxelPositionalEmbedding, sinusoid_table, voxel_pe
    m = VoxelPositionalEmbedding().cuda()
    assert m.pos_embedding.shape == (256, 469) and m.pos_embedding.is_cuda and len(m.state_dict()) == 0
    feat = torch.randn(2, 64, 1408, device="cuda")
    pc = torch.randint(0, 256, (2, 64, 3), device="cuda").float()
    out = m(feat, pc)
    want = orc.voxel_pe(feat.cpu().numpy(), pc.cpu().numpy(), sinusoid_table().numpy())
    assert np.array_equal(out.cpu().numpy(), want)
    assert torch.equal(out[..., 1407], feat[..., 1407])                      # the 1408th channel gets no embedding
    assert m(feat, pc, mode="cat").shape == (2, 128, 1408)
    assert voxel_pe(feat[:0], pc[:0], m.pos_embedding).shape == (0, 64, 1408)
    for bad in (256.0, -257.0, float("nan"), float("inf")):
        pc2 = pc.clone()
        pc2[1, 63, 2] = bad
        with pytest.raises(IndexError):
            m(feat, pc2)
    _, status = voxel_pe(feat, pc, m.pos_embedding, validate=False)
    assert int(status.item()) == 0
    with pytest.raises(RuntimeError):
        voxel_pe(feat.cpu(), pc.cpu(), m.pos_embedding.cpu())                  # no CPU path
    with pytest.raises(RuntimeError):
        voxel_pe(feat[..., :1000].contiguous(), pc, m.pos_embedding)            # 3 * 469 channels do not fit


def _projection_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_projection.npz"))


def test_projection_vs_reference_class():
    """csrc/projection.cu behind the reference's ProjectionHelper interface against the outputs of the reference's own
    class (lib/projection.py run unmodified on the CPU by tests/golden/make_ref_projection_goldens.py): the index
    lists of every view are identical (batched and one view at a time, `None` for the view that sees nothing),
    project() is identical, corners / normals agree to fp32 rounding."""
    from situation3d_b200.projection import ProjectionHelper
    g = _projection_golden()
    dmin, dmax, acc = (float(v) for v in g["params"])
    helper = ProjectionHelper(torch.from_numpy(g["intrinsic"]), dmin, dmax, [int(v) for v in g["image_dims"]], acc)
    assert np.array_equal(helper.corner_points[:, :3].cpu().numpy(), g["corner_points"]) and helper.corner_points.is_cuda
    points = torch.from_numpy(g["points"]).cuda()
    poses, depths = torch.from_numpy(g["poses"]).cuda(), torch.from_numpy(g["depths"]).cuda()
    n, V = points.shape[0], poses.shape[0]
    i3, i2, counts = helper.compute_projection_views(points, depths, poses)
    assert i3.dtype == torch.int64 and i3.shape == (V, n + 1)
    assert np.array_equal(i3.cpu().numpy(), g["indices_3d"].astype(np.int64))
    assert np.array_equal(i2.cpu().numpy(), g["indices_2d"].astype(np.int64))
    assert np.array_equal(counts.cpu().numpy(), g["indices_3d"][:, 0])
    for v in range(V):
        res = helper.compute_projection(points, depths[v], poses[v])
        if g["indices_3d"][v, 0] == 0:
            assert res is None
        else:
            assert np.array_equal(res[0].cpu().numpy(), g["indices_3d"][v]) and np.array_equal(res[1].cpu().numpy(), g["indices_2d"][v])
        cc = helper.compute_frustum_corners(poses[v])
        assert cc.shape == (8, 4, 1)
        np.testing.assert_allclose(cc.squeeze(-1).cpu().numpy(), g["corners"][v], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(helper.compute_frustum_normals(cc).cpu().numpy(), g["normals"][v], rtol=1e-4, atol=1e-4)
        # points_in_frustum on the reference's own corners / normals: mask and count
        want = np.unpackbits(g["frustum_masks"][v])[:n].astype(bool)
        rc, rn = torch.from_numpy(g["corners"][v]).cuda().unsqueeze(2), torch.from_numpy(g["normals"][v]).cuda()
        mask = helper.points_in_frustum(rc, rn, points.cpu(), return_mask=True)
        assert mask.dtype == torch.bool and np.array_equal(mask.cpu().numpy(), want)
        assert int(helper.points_in_frustum(rc, rn, points)) == int(want.sum())
    label = torch.from_numpy(g["label"]).cuda()
    from situation3d_b200.projection import Projection
    assert torch.equal(Projection.apply(label[3], i3[3], i2[3], n), helper.project(label[3], i3[3], i2[3], n))
    out = helper.project(label[3], i3[3], i2[3], n)
    assert out.shape == (label.shape[1], n) and np.array_equal(out.cpu().numpy(), g["project_view3"])
    allv = helper.project_views(label, i3, i2, n)
    assert torch.equal(allv[3], out)
    np.testing.assert_allclose(allv.double().sum((1, 2)).cpu().numpy(), g["project_checksum"], rtol=1e-9, atol=1e-9)
    # across-view pooling = the running max over the reference's per-frame project() outputs, from zeros
    pooled = helper.project_views_maxpool(label, i3, i2, n)
    want = torch.zeros_like(allv[0])
    for v in range(allv.shape[0]):
        want = torch.maximum(want, allv[v])
    assert torch.equal(pooled, want.t().contiguous())
    assert torch.equal(helper.project_views_maxpool(label, i3, i2, n, layout="channels"), want)
    g3, g2 = torch.from_numpy(g["indices_3d"]).cuda(), torch.from_numpy(g["indices_2d"]).cuda()
    ref3 = torch.from_numpy(g["project_view3"]).cuda()                       # the reference's own project() of view 3
    only3 = helper.project_views_maxpool(label[3:4], g3[3:4], g2[3:4], n, layout="channels")
    assert torch.equal(only3, torch.clamp_min(ref3, 0.0))
    one = helper.project(label[3, 0], i3[3], i2[3], n)                       # a single (H, W) plane -> (1, n)
    assert one.shape == (1, n) and torch.equal(one[0], out[0])


@pytest.mark.parametrize("n,V,C", [(50000, 24, 128), (4097, 3, 5), (1, 2, 1), (300, 1, 33)])
def test_projection_vs_oracle(n, V, C):
    """Scene-size clouds and ragged shapes (segment boundaries at 4096 points) against the C oracle, bit for bit;
    then the properties the lists must have (ascending points, count first, zero padding) and the back-projection
    against numpy."""
    from situation3d_b200.projection import ProjectionHelper
    from situation3d_b200.synthetic import make_scene, make_views
    pts = make_scene(5, n_points=max(n, 64), n_features=1)[:n, :3].astype(np.float32)
    intrinsic, poses, depths = make_views(pts, V, seed=n)
    dims, dmin, dmax, acc = [41, 32], 0.4, 4.0, 0.05
    helper = ProjectionHelper(torch.from_numpy(intrinsic), dmin, dmax, dims, acc)
    points = torch.from_numpy(pts).cuda()
    c2w = torch.from_numpy(poses).cuda()
    w2c = torch.inverse(c2w)
    i3, i2, counts = helper.compute_projection_views(points, torch.from_numpy(depths).cuda(), c2w, w2c)
    i3h, i2h, w2ch = i3.cpu(), i2.cpu(), w2c.cpu()
    total = 0
    for v in range(V):
        want = orc.compute_projection(torch.from_numpy(pts), torch.from_numpy(depths[v]), torch.from_numpy(poses[v]), w2ch[v],
                                      torch.from_numpy(intrinsic), dmin, dmax, dims, acc)
        k = int(i3h[v, 0])
        if want is None:
            assert k == 0 and int(i3h[v].abs().sum()) == 0 and int(i2h[v].abs().sum()) == 0
            continue
        assert torch.equal(i3h[v], want[0]) and torch.equal(i2h[v], want[1]), v
        assert k == int(counts[v]) and bool((i3h[v, 2:1 + k] > i3h[v, 1:k]).all()) and int(i3h[v, 1 + k:].abs().sum()) == 0
        total += k
    if n >= 4097:
        assert total > 0                                                       # the synthetic views do see the cloud
    label = torch.randn(V, C, dims[1], dims[0], generator=torch.Generator().manual_seed(n))
    got = helper.project_views(label.cuda(), i3, i2, n).cpu().numpy()
    for v in range(V):
        assert np.array_equal(got[v], orc.project(label[v].numpy(), i3h[v].numpy(), i2h[v].numpy(), n)), v


def test_projection_errors():
    from situation3d_b200.projection import ProjectionHelper
    helper = ProjectionHelper(torch.eye(4), 0.4, 4.0, [41, 32], 0.05)
    with pytest.raises(RuntimeError):
        ProjectionHelper(torch.eye(4), 0.4, 4.0, [41, 32], 0.05, cuda=False)          # no CPU path
    with pytest.raises(RuntimeError):
        helper.compute_projection(torch.zeros(10, 3), torch.zeros(32, 41), torch.eye(4))
    i3 = torch.zeros(11, dtype=torch.int64, device="cuda")
    i2 = torch.zeros(11, dtype=torch.int64, device="cuda")
    i3[0], i2[0], i3[1], i2[1] = 1, 1, 4, 32 * 41                                      # pixel index one past the image
    with pytest.raises(IndexError):
        helper.project(torch.zeros(2, 32, 41, device="cuda"), i3, i2, 10)
    i2[1] = 5
    lab = torch.arange(2 * 32 * 41, dtype=torch.float32, device="cuda").reshape(2, 32, 41)
    out = helper.project(lab, i3, i2, 10)
    assert out[:, 4].tolist() == [5.0, 5.0 + 32 * 41] and float(out.abs().sum()) == 10.0 + 32 * 41
