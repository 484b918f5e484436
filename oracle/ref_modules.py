"""The reference's own Python modules (lib/pointnet2/{pytorch_utils,pointnet2_utils,pointnet2_modules}.py),
UNMODIFIED, made loadable on the GPU box.  TEST INFRASTRUCTURE ONLY.

/root/reference does not exist on the GPU box and its sources are never copied into this repository.  Like the
reference's CUDA extension (oracle/build_ref.py -> oracle/_ref/pn2_ref_ext.so), the three modules are therefore
COMPILED where they lie -- ``py_compile`` to byte code -- with the outputs only in oracle/_ref/ (git-ignored, travels
with the gpurun snapshot).  ``load(ext)`` imports that byte code with ``ext`` (the reference extension, or any module
with the nine ``_ext`` functions) registered as ``pointnet2._ext``, exactly what pointnet2_utils.py:25-33 imports.

``RefBackbone`` is the Pointnet2Backbone composition (absent from the release) over the
reference's PointnetSAModuleVotes / PointnetFPModule classes -- bench.py's ``reference_cuda`` arm.
"""
import importlib.machinery
import importlib.util
import os
import py_compile
import sys
import types

REF_PY = "/root/reference/lib/pointnet2"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
NAMES = ("pytorch_utils", "pointnet2_utils", "pointnet2_modules")


def pyc_path(name):
    # (not *.pyc: snapshot tools tend to drop those as caches)
    return os.path.join(OUT_DIR, name + ".bytecode")


def build():
    """Byte-compile the three reference modules into oracle/_ref/ (no-op without /root/reference)."""
    if not os.path.isdir(REF_PY):
        return all(os.path.exists(pyc_path(n)) for n in NAMES)
    os.makedirs(OUT_DIR, exist_ok=True)
    for n in NAMES:
        src = os.path.join(REF_PY, n + ".py")
        if not os.path.exists(pyc_path(n)) or os.path.getmtime(pyc_path(n)) < os.path.getmtime(src):
            py_compile.compile(src, cfile=pyc_path(n), doraise=True)
    return True


def available():
    return all(os.path.exists(pyc_path(n)) for n in NAMES)


def load(ext):
    """Returns the reference's ``pointnet2_modules`` module running over ``ext``; None when the byte code is absent."""
    if not available():
        return None
    pkg = types.ModuleType("pointnet2")
    pkg._ext = ext
    sys.modules["pointnet2"] = pkg
    sys.modules["pointnet2._ext"] = ext
    mods = {}
    for n in NAMES:          # dependency order: the later modules import the earlier ones by bare name
        loader = importlib.machinery.SourcelessFileLoader(n, pyc_path(n))
        spec = importlib.util.spec_from_loader(n, loader)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[n] = mod
        loader.exec_module(mod)
        mods[n] = mod
    return mods["pointnet2_modules"]


def make_backbone(ref_modules, input_feature_dim=129):
    """Pointnet2Backbone over the reference's module classes."""
    import torch.nn as nn

    class RefBackbone(nn.Module):
        def __init__(self):
            super().__init__()
            SA, FP = ref_modules.PointnetSAModuleVotes, ref_modules.PointnetFPModule
            kw = dict(use_xyz=True, normalize_xyz=True)
            self.sa1 = SA(npoint=2048, radius=0.2, nsample=64, mlp=[input_feature_dim, 64, 64, 128], **kw)
            self.sa2 = SA(npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256], **kw)
            self.sa3 = SA(npoint=512, radius=0.8, nsample=16, mlp=[256, 128, 128, 256], **kw)
            self.sa4 = SA(npoint=256, radius=1.2, nsample=16, mlp=[256, 128, 128, 256], **kw)
            self.fp1 = FP(mlp=[256 + 256, 256, 256])
            self.fp2 = FP(mlp=[256 + 256, 256, 256])

        def forward(self, pc):
            xyz = pc[..., 0:3].contiguous()
            features = pc[..., 3:].transpose(1, 2).contiguous()
            out = {}
            for lvl, m in enumerate((self.sa1, self.sa2, self.sa3, self.sa4), start=1):
                xyz, features, inds = m(xyz, features)
                out["sa%d_xyz" % lvl], out["sa%d_features" % lvl], out["sa%d_inds" % lvl] = xyz, features, inds
            f = self.fp1(out["sa3_xyz"], out["sa4_xyz"], out["sa3_features"], out["sa4_features"])
            f = self.fp2(out["sa2_xyz"], out["sa3_xyz"], out["sa2_features"], f)
            out["fp2_features"], out["fp2_xyz"] = f, out["sa2_xyz"]
            out["fp2_inds"] = out["sa1_inds"][:, 0:out["fp2_xyz"].shape[1]]
            return out

    return RefBackbone()


if __name__ == "__main__":
    print("reference modules byte-compiled:", build())
