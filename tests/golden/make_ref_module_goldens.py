"""Golden vectors from the REFERENCE's own Python modules (runs in the build container only,
where /root/reference exists; the GPU box never reads it).

    python tests/golden/make_ref_module_goldens.py        # writes tests/golden/ref_modules.npz

The reference's lib/pointnet2/{pointnet2_modules,pointnet2_utils,pytorch_utils}.py are imported
unmodified from /root/reference.  Their native dependency ``pointnet2._ext`` only exists for
CUDA, so a stand-in module with the same nine function names is registered that executes the
CPU oracle (oracle/pn2_oracle.c) -- itself pinned to the reference CUDA kernels by
ref_cuda_ops.npz.  The situation helpers are taken from the reference source by compiling the
individual function definitions out of situation3d/models/sqa_module.py and
situation3d/utils/temp.py (those files cannot be imported whole: sqa_module pulls in
MinkowskiEngine and a dataset-scanning config, temp.py is a scratch file missing its imports).
Nothing from the reference is copied into the repository; only inputs and outputs are stored.
"""
import ast
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import pn2_oracle as orc  # noqa: E402


def install_ext_stub():
    pkg = types.ModuleType("pointnet2")
    ext = types.ModuleType("pointnet2._ext")
    for name in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
                 "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"):
        setattr(ext, name, getattr(orc, name))
    pkg._ext = ext
    sys.modules["pointnet2"] = pkg
    sys.modules["pointnet2._ext"] = ext


def functions_from(path, names):
    """Compile only the named top-level function definitions of a reference source file."""
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(body) == len(names), (path, names)
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def sd_to_np(prefix, module, out):
    for k, v in module.state_dict().items():
        out[prefix + k] = v.numpy()


def randomize_bn(module, g):
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))


def main():
    install_ext_stub()
    sys.path.insert(0, os.path.join(REF, "lib", "pointnet2"))
    import pointnet2_modules as ref_modules   # the reference's file

    g = torch.Generator().manual_seed(4242)
    torch.manual_seed(4242)
    out = {}

    # PointnetSAModuleVotes: the backbone configuration at reduced size, and a variant without normalisation
    cases = [("sa_a", dict(npoint=32, radius=0.6, nsample=16, mlp=[6, 16, 16, 32], use_xyz=True, normalize_xyz=True)),
             ("sa_b", dict(npoint=20, radius=0.9, nsample=8, mlp=[6, 24, 40], use_xyz=True, normalize_xyz=False)),
             ("sa_c", dict(npoint=16, radius=0.7, nsample=80, mlp=[6, 16, 32], use_xyz=True, normalize_xyz=True))]
    for name, kw in cases:
        mod = ref_modules.PointnetSAModuleVotes(**{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
        randomize_bn(mod, g)
        mod.eval()
        xyz = torch.randn(2, 400, 3, generator=g)
        feats = torch.randn(2, 6, 400, generator=g)
        with torch.no_grad():
            new_xyz, new_feats, inds = mod(xyz, feats)
        out[name + "_xyz"], out[name + "_feats"] = xyz.numpy(), feats.numpy()
        out[name + "_new_xyz"], out[name + "_new_feats"], out[name + "_inds"] = \
            new_xyz.numpy(), new_feats.numpy(), inds.numpy()
        sd_to_np(name + "_sd_", mod, out)

    # PointnetFPModule
    mod = ref_modules.PointnetFPModule(mlp=[32 + 12, 48, 24])
    randomize_bn(mod, g)
    mod.eval()
    unknown, known = torch.randn(2, 90, 3, generator=g), torch.randn(2, 30, 3, generator=g)
    uf, kf = torch.randn(2, 12, 90, generator=g), torch.randn(2, 32, 30, generator=g)
    with torch.no_grad():
        res = mod(unknown, known, uf, kf)
    out["fp_unknown"], out["fp_known"], out["fp_uf"], out["fp_kf"], out["fp_out"] = \
        unknown.numpy(), known.numpy(), uf.numpy(), kf.numpy(), res.numpy()
    sd_to_np("fp_sd_", mod, out)

    # QueryAndGroup alone (grouped tensor incl. xyz-first channel order)
    import pointnet2_utils as ref_utils
    grouper = ref_utils.QueryAndGroup(0.8, 12, use_xyz=True, ret_grouped_xyz=True, normalize_xyz=True)
    xyz = torch.randn(1, 200, 3, generator=g)
    new_xyz = xyz[:, :10].contiguous()
    feats = torch.randn(1, 4, 200, generator=g)
    nf, gx = grouper(xyz, new_xyz, feats)
    out["qg_xyz"], out["qg_new_xyz"], out["qg_feats"], out["qg_out"], out["qg_gxyz"] = \
        xyz.numpy(), new_xyz.numpy(), feats.numpy(), nf.numpy(), gx.numpy()

    # situation helpers: the reference's own function bodies
    q2r, rv2m = functions_from(os.path.join(REF, "situation3d", "models", "sqa_module.py"),
                               ["quaternions_to_rotation_matrices", "batch_rotation_vector_to_matrix"])
    (bmf,) = functions_from(os.path.join(REF, "situation3d", "utils", "temp.py"), ["batch_matrix_function"])
    quats = torch.randn(6, 4, generator=g)
    out["sit_quats"], out["sit_quat_R"] = quats.numpy(), q2r(quats).numpy()
    rv = torch.randn(6, 3, generator=g)
    rv[1] = 0.0
    rv[4] = 1e-8
    out["sit_rotvec"], out["sit_rotvec_R"] = rv.numpy(), rv2m(rv).numpy()
    sit = torch.randn(6, 7, generator=g)
    M = bmf(sit)
    out["sit_vec"], out["sit_M"] = sit.numpy(), M.numpy()
    # temp.py:86-97 transform of (B,256,3) tokens, then SIG3D.pos_embed (sqa_module.py:274-278,319-321)
    # and the Gaussian prior (sqa_module.py:328-336), evaluated with the same torch statements
    pos = torch.randn(6, 40, 3, generator=g)
    aug = torch.cat([pos, torch.ones(6, 40, 1)], dim=2)
    out["sit_pos"] = pos.numpy()
    out["sit_pos_t"] = torch.bmm(aug, M.transpose(1, 2))[:, :, :3].numpy()
    pe = torch.nn.Sequential(torch.nn.Linear(2, 128), torch.nn.GELU(), torch.nn.Linear(128, 256))
    tokens = torch.randn(6, 40, 256, generator=g)
    with torch.no_grad():
        out["sit_tokens"] = tokens.numpy()
        out["sit_tokens_pe"] = (tokens + pe(pos[..., :2])).numpy()
        d = torch.norm(pos[..., :2] - sit[:, :3].unsqueeze(1)[:, :, :2], dim=2)
        w = torch.exp(-d ** 2 / (2 * 0.16 ** 2))
        out["sit_prior"] = (w / w.sum(dim=1, keepdim=True)).numpy()
    sd_to_np("sit_pe_", pe, out)

    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_modules.npz"), **out)
    print("wrote ref_modules.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
