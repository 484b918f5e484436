"""Build the reference's own CUDA extension, unmodified, into oracle/_ref/.

TEST INFRASTRUCTURE ONLY (see oracle/pn2_oracle.c).  The sources are compiled
where they lie under /root/reference/lib/pointnet2/_ext_src (nothing is copied
into this repository); only the resulting shared object lands in oracle/_ref/,
which is git-ignored but travels to the GPU box with the gpurun snapshot.

The reference's setup.py cannot be used as is: it hard-codes
TORCH_CUDA_ARCH_LIST="3.7+PTX;...;7.5" (setup.py:17), none of which nvcc 12.9
still knows and none of which runs on a B200.  The same nine source files are
therefore handed to torch.utils.cpp_extension.load() with the arch list set to
10.0 and the reference's own flags (setup.py:29-32: -O3 for both compilers).

The module is named ``pn2_ref_ext`` (not ``pointnet2._ext``) so that it can
coexist with this repository's own drop-in ``pointnet2._ext`` alias.
"""
import glob
import os
import sys

REF_SRC = "/root/reference/lib/pointnet2/_ext_src"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
MODULE = "pn2_ref_ext"


def so_path():
    return os.path.join(OUT_DIR, MODULE + ".so")


def build(verbose=False):
    """Returns the path of the built module, or None when the reference tree is absent."""
    if not os.path.isdir(REF_SRC):
        return so_path() if os.path.exists(so_path()) else None
    sources = sorted(glob.glob(os.path.join(REF_SRC, "src", "*.cpp")) +
                     glob.glob(os.path.join(REF_SRC, "src", "*.cu")))
    newest = max(os.path.getmtime(s) for s in sources)
    if os.path.exists(so_path()) and os.path.getmtime(so_path()) >= newest:
        return so_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils.cpp_extension import load
    load(name=MODULE, sources=sources,
         extra_include_paths=[os.path.join(REF_SRC, "include")],
         extra_cflags=["-O3"], extra_cuda_cflags=["-O3"],
         build_directory=OUT_DIR, verbose=verbose, is_python_module=False)
    return so_path()


def load_ref_ext():
    """Import oracle/_ref/pn2_ref_ext.so (prebuilt).  Returns None if it is not there."""
    path = so_path()
    if not os.path.exists(path):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location(MODULE, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print("reference extension:", p)
