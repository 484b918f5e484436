"""CPU suite, part 2: the C-ABI library loads and exports what include/pn2_b200.h declares, and
the host-side mirror of the reference interface behaves like the reference (no compute calls)."""
import ctypes
import inspect
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pn2_b200.h")
REF_PN2 = "/root/reference/lib/pointnet2"


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pn2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from situation3d_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(_lib.lib, name), "libpn2_b200.so does not export %s" % name
        assert name in _lib._PROTOTYPES, "%s has no ctypes prototype" % name
    assert set(_lib._PROTOTYPES) <= set(names), "prototype without a declaration in the header"


def test_library_basics_without_gpu():
    from situation3d_b200 import _lib
    assert _lib.lib.pn2_version() >= 100
    assert _lib.lib.pn2_error_string(0) == b"ok"
    assert b"invalid" in _lib.lib.pn2_error_string(-1)
    dims = _lib.int_array([132, 64, 64, 128])
    assert _lib.lib.pn2_mlp_f32_image_bytes(3, dims) == 4 * (132 * 64 + 64 + 64 * 64 + 64 + 64 * 128 + 128)
    assert _lib.lib.pn2_mlp_f32_supported(3, dims) == 1
    assert _lib.lib.pn2_mlp_f32_supported(2, _lib.int_array([4096, 4096, 64])) == 0
    # large scenes: the bucketed kernel's scratch, per scene 20 bytes per (padded) point + 2 per point, 256-byte aligned
    # large scenes: the bucketed kernel's scratch, per scene 20 bytes per point (padded to whole 128-point buckets) + 2
    # per point, 256-byte aligned
    assert _lib.lib.pn2_furthest_point_sampling_workspace_bytes(8, 200000, 4096) == 8 * ((20 * 200064 + 2 * 200000 + 32 + 4 * 32768 + 255) // 256 * 256)
    assert _lib.lib.pn2_furthest_point_sampling_workspace_bytes(8, 40000, 2048) == 0   # register-resident cluster kernel
    assert _lib.lib.pn2_furthest_point_sampling_workspace_bytes(8, 2048, 1024) == 0
    # PN2_FPS_THROUGHPUT (1): the bucketed kernel already from 32 768 points; PN2_FPS_LATENCY (0) = the plain query
    assert _lib.lib.pn2_furthest_point_sampling_workspace_bytes_mode(8, 40000, 2048, 0) == 0
    assert _lib.lib.pn2_furthest_point_sampling_workspace_bytes_mode(8, 40000, 2048, 1) > 0
    assert _lib.lib.pn2_furthest_point_sampling_workspace_bytes_mode(8, 20000, 2048, 1) == 0
    assert _lib.lib.pn2_furthest_point_sampling_workspace_bytes_mode(8, 40000, 2048, 7) == 0
    # argument validation happens before any CUDA call
    assert _lib.lib.pn2_gather_points(-1, 1, 1, 1, None, None, None, None) == -1
    assert _lib.lib.pn2_ball_query(1, 10, 4, 0.5, 8, None, None, None, None) == -1


def test_ext_rejects_cpu_tensors_like_the_reference():
    # utils.h:5-25: AT_ASSERT -> RuntimeError for CPU / wrong dtype / non-contiguous tensors
    from situation3d_b200.pointnet2 import _ext
    x = torch.zeros(1, 8, 3)
    with pytest.raises(RuntimeError):
        _ext.furthest_point_sampling(x, 4)
    with pytest.raises(RuntimeError):
        _ext.ball_query(x, x, 0.5, 4)
    with pytest.raises(RuntimeError):
        _ext.gather_points(torch.zeros(1, 3, 8), torch.zeros(1, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError):
        _ext.three_nn(x.double(), x)
    assert sorted(_ext.__all__) == sorted([
        "gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
        "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"])   # bindings.cpp:6-19


def test_backbone_state_dict_keys():
    # keys a reference/VoteNet checkpoint carries
    from situation3d_b200.backbone_module import Pointnet2Backbone
    net = Pointnet2Backbone(input_feature_dim=129)
    keys = set(net.state_dict().keys())
    for sa in ("sa1", "sa2", "sa3", "sa4"):
        for i in range(3):
            assert "%s.mlp_module.layer%d.conv.weight" % (sa, i) in keys
            for leaf in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
                assert "%s.mlp_module.layer%d.bn.bn.%s" % (sa, i, leaf) in keys
    for fp in ("fp1", "fp2"):
        for i in range(2):
            assert "%s.mlp.layer%d.conv.weight" % (fp, i) in keys
            assert "%s.mlp.layer%d.bn.bn.running_var" % (fp, i) in keys
    assert not any("conv.bias" in k for k in keys)          # bias only without BN (pytorch_utils.py:86)
    assert net.sa1.mlp_module.layer0.conv.weight.shape == (64, 132, 1, 1)
    assert net.sa2.mlp_module.layer0.conv.weight.shape == (128, 131, 1, 1)
    assert net.fp2.mlp.layer1.conv.weight.shape == (256, 256, 1, 1)
    assert len(keys) == 4 * 3 * 6 + 2 * 2 * 6


def test_mlp_spec_is_modified_in_place_like_the_reference():
    from situation3d_b200.pointnet2.pointnet2_modules import PointnetSAModuleVotes
    spec = [6, 16, 32]
    PointnetSAModuleVotes(mlp=spec, npoint=4, radius=0.5, nsample=8)
    assert spec[0] == 9                                       # pointnet2_modules.py:204-206


def test_bn_folding_equals_eval_forward():
    from situation3d_b200.fused import fold_shared_mlp
    from situation3d_b200.pointnet2.pytorch_utils import SharedMLP
    from situation3d_b200.synthetic import randomize_bn_stats
    torch.manual_seed(0)
    mlp = randomize_bn_stats(SharedMLP([7, 12, 9], bn=True)).eval()
    x = torch.randn(2, 7, 5, 3)
    want = mlp(x)
    y = x
    for w, b in fold_shared_mlp(mlp):
        y = torch.relu(torch.einsum("oc,bcps->bops", w, y) + b[None, :, None, None])
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-5)
    assert fold_shared_mlp(SharedMLP([4, 4], bn=True, activation=torch.nn.Tanh())) is None
    assert fold_shared_mlp(SharedMLP([4, 4], bn=True, preact=True)) is None


def test_loads_reference_state_dict(ref_modules_golden):
    # a state_dict produced by the reference's own PointnetSAModuleVotes loads key for key
    from situation3d_b200.pointnet2.pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes
    g = ref_modules_golden
    sd = {k[len("sa_a_sd_"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sa_a_sd_")}
    mod = PointnetSAModuleVotes(npoint=32, radius=0.6, nsample=16, mlp=[6, 16, 16, 32], normalize_xyz=True)
    assert set(mod.state_dict().keys()) == set(sd.keys())
    mod.load_state_dict(sd, strict=True)
    sd = {k[len("fp_sd_"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("fp_sd_")}
    PointnetFPModule(mlp=[44, 48, 24]).load_state_dict(sd, strict=True)


@pytest.mark.skipif(not os.path.isdir(REF_PN2), reason="reference tree not present")
def test_signatures_match_reference():
    """Same public names, constructor and forward parameters as lib/pointnet2 (extra keyword-only
    arguments with defaults are allowed: fused, precision)."""
    import importlib
    import types
    from oracle import pn2_oracle as orc
    saved = {k: sys.modules.get(k) for k in ("pointnet2", "pointnet2._ext", "pointnet2_utils", "pytorch_utils",
                                             "pointnet2_modules")}
    pkg, ext = types.ModuleType("pointnet2"), types.ModuleType("pointnet2._ext")
    for name in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
                 "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"):
        setattr(ext, name, getattr(orc, name))
    pkg._ext = ext
    sys.modules["pointnet2"], sys.modules["pointnet2._ext"] = pkg, ext
    sys.path.insert(0, REF_PN2)
    try:
        ref_mod = importlib.import_module("pointnet2_modules")
        ref_utils = importlib.import_module("pointnet2_utils")
        ref_pt = importlib.import_module("pytorch_utils")
    finally:
        sys.path.remove(REF_PN2)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    from situation3d_b200.pointnet2 import pointnet2_modules as my_mod, pointnet2_utils as my_utils, \
        pytorch_utils as my_pt

    def params(fn):
        return [(p.name, p.kind, p.default if not callable(p.default) and not isinstance(p.default, torch.nn.Module)
                 else "callable") for p in inspect.signature(fn).parameters.values()]

    pairs = [(ref_mod, my_mod, ["PointnetSAModuleVotes", "PointnetFPModule", "PointnetSAModuleMSG", "PointnetSAModule",
                                "PointnetSAModuleMSGVotes", "PointnetLFPModuleMSG"]),
             (ref_utils, my_utils, ["QueryAndGroup", "GroupAll", "RandomDropout"]),
             (ref_pt, my_pt, ["SharedMLP", "Conv1d", "Conv2d", "Conv3d", "BatchNorm1d", "BatchNorm2d", "BatchNorm3d",
                              "FC", "BNMomentumScheduler"])]
    for ref, mine, names in pairs:
        for name in names:
            r, m = getattr(ref, name), getattr(mine, name)
            rp, mp = params(r.__init__), params(m.__init__)
            assert mp[:len(rp)] == rp, (name, rp, mp)
            assert all(p[0] in ("fused", "precision") for p in mp[len(rp):]), name
            if hasattr(r, "forward") and name not in ("SharedMLP", "Conv1d", "Conv2d", "Conv3d", "BatchNorm1d",
                                                      "BatchNorm2d", "BatchNorm3d", "FC"):
                assert [p[0] for p in params(m.forward)] == [p[0] for p in params(r.forward)], name
    for fn in ("furthest_point_sample", "gather_operation", "three_nn", "three_interpolate", "grouping_operation",
               "ball_query"):
        assert hasattr(my_utils, fn)
    # identical parameter names/shapes for an identically configured module
    a = ref_mod.PointnetSAModuleVotes(npoint=8, radius=0.3, nsample=4, mlp=[5, 8, 16])
    b = my_mod.PointnetSAModuleVotes(npoint=8, radius=0.3, nsample=4, mlp=[5, 8, 16])
    assert {k: tuple(v.shape) for k, v in a.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in b.state_dict().items()}


def test_split_first_layer_rule():
    """Host logic of the split first SA layer: taken only where it shrinks the fused kernel's layer-1 K."""
    from situation3d_b200 import fused
    assert fused.split_first_layer(129, 64)        # SA1: K 144 -> 80
    assert not fused.split_first_layer(128, 128)   # SA2: 144 -> 144
    assert fused.split_first_layer(256, 128)       # SA3 / SA4: 272 -> 144
    assert not fused.split_first_layer(129, 96)    # widths the per-point GEMM does not cover
    assert not fused.split_first_layer(6, 64)


REF_PROJECTION = "/root/reference/lib/projection.py"


@pytest.mark.skipif(not os.path.isfile(REF_PROJECTION), reason="reference tree not present")
def test_projection_helper_signatures_match_reference():
    """ProjectionHelper keeps the reference's constructor and method signatures (lib/projection.py:5-279); the batched
    forms are additions."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_projection_sig", REF_PROJECTION)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from situation3d_b200.projection import ProjectionHelper

    def params(fn):
        return [(p.name, p.kind, p.default) for p in inspect.signature(fn).parameters.values()]

    for name in ("__init__", "depth_to_skeleton", "skeleton_to_depth", "compute_frustum_corners", "compute_frustum_normals",
                 "points_in_frustum", "compute_projection", "project"):
        assert params(getattr(ProjectionHelper, name)) == params(getattr(ref.ProjectionHelper, name)), name
    assert hasattr(ProjectionHelper, "compute_projection_views") and hasattr(ProjectionHelper, "project_views")
    from situation3d_b200.projection import Projection
    assert params(Projection.forward) == params(ref.Projection.forward)


def test_voxel_pe_and_projection_reject_cpu_tensors():
    """The widened rows have no CPU path either: host tensors raise instead of being computed somewhere else."""
    from situation3d_b200.voxel_pe import voxel_pe
    with pytest.raises(RuntimeError):
        voxel_pe(torch.zeros(1, 2, 1408), torch.zeros(1, 2, 3), torch.zeros(4, 469))
    from situation3d_b200.projection import ProjectionHelper
    with pytest.raises(RuntimeError):
        ProjectionHelper(torch.eye(4), 0.4, 4.0, [41, 32], 0.05, cuda=False)


def test_reencoder_training_path_matches_oracle_and_differentiates():
    """SituationReencoder in train mode (or whenever gradients are needed) runs the PyTorch twin of the fused kernel:
    same numbers as the oracle, and pos_embed / tokens receive gradients (the fused kernel is forward-only)."""
    import torch
    from oracle import pn2_oracle as orc
    from situation3d_b200.reencode import SituationReencoder
    from situation3d_b200.synthetic import make_situations
    torch.manual_seed(0)
    for agent in (False, True):
        re = SituationReencoder(to_agent_frame=agent).train()
        tokens = torch.randn(3, 40, 256, requires_grad=True)
        pos = torch.randn(3, 40, 3)
        sit = torch.from_numpy(make_situations(3))
        d = re({"scene_feat": tokens, "scene_positions": pos, "auxiliary_task": sit})
        pe = re.pos_embed
        want_tok, want_pos, want_prior = orc.reencode(tokens.detach(), pos, sit, pe[0].weight.detach(), pe[0].bias.detach(),
                                                      pe[2].weight.detach(), pe[2].bias.detach(), to_agent_frame=agent)
        torch.testing.assert_close(d["scene_feat"], want_tok, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(d["scene_positions_agent"], want_pos, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(d["auxiliary_task_loc_gt"], want_prior, rtol=1e-4, atol=1e-8)
        d["scene_feat"].square().mean().backward()
        assert tokens.grad is not None and pe[0].weight.grad is not None and pe[2].weight.grad.abs().sum() > 0


def test_sampling_mode_switch_and_training_fallbacks_on_cpu():
    """Host logic that needs no GPU: the sampling-mode context manager restores the previous mode (also on error), the
    pipeline's rule of thumb, and the row-layout training helpers fall back to PyTorch ops for CPU tensors."""
    import torch
    from situation3d_b200 import fused, train_rows
    from situation3d_b200.graphs import pick_sampling_mode
    assert fused._SAMPLING_MODE[0] == 0
    with fused.sampling_mode("throughput"):
        assert fused._SAMPLING_MODE[0] == 1
        with fused.sampling_mode("latency"):
            assert fused._SAMPLING_MODE[0] == 0
        assert fused._SAMPLING_MODE[0] == 1
    assert fused._SAMPLING_MODE[0] == 0
    with pytest.raises(KeyError):
        with fused.sampling_mode("fastest"):
            pass
    with pytest.raises(ZeroDivisionError):
        with fused.sampling_mode("throughput"):
            1 / 0
    assert fused._SAMPLING_MODE[0] == 0
    assert pick_sampling_mode(8) == "latency" and pick_sampling_mode(127) == "latency" and pick_sampling_mode(128) == "throughput"
    # SharedMLP on rows with CPU tensors: the hand-written BatchNorm kernels are refused, torch ops give the module's result
    from situation3d_b200.pointnet2 import pytorch_utils as pt_utils
    torch.manual_seed(0)
    mlp = pt_utils.SharedMLP([7, 16, 8], bn=True).train()
    x = torch.randn(2 * 5 * 4, 7)
    assert not train_rows._fused_bn_relu_ok(mlp[0], x, 0)
    assert not train_rows._group_rows_ok(torch.randn(2, 10, 4), torch.randn(2, 10, 3))
    import copy
    ref = copy.deepcopy(mlp)
    got = train_rows.shared_mlp_rows(mlp, x, pool_ns=4)                                      # (10, 8)
    want = ref(x.view(2, 5, 4, 7).permute(0, 3, 1, 2)).amax(dim=3).permute(0, 2, 1).reshape(10, 8)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    for (n1, b1), (_, b2) in zip(mlp.named_buffers(), ref.named_buffers()):
        torch.testing.assert_close(b1.float(), b2.float(), rtol=1e-5, atol=1e-6, msg=n1)
