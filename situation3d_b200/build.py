"""Build libpn2_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m situation3d_b200.build [-v] [--force]

The shared object lands next to this file (situation3d_b200/libpn2_b200.so); it is
git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles
without a GPU, so this also runs in the CPU-only build container.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PN2_BUILD_SUFFIX / PN2_BUILD_DEFINES: a second library next to the default one, compiled with extra -D switches, for
# A/B measurements on the GPU box (loaded through PN2_B200_LIB, see _lib.py); the default build ignores both
SUFFIX = os.environ.get("PN2_BUILD_SUFFIX", "")
LIB = os.path.join(HERE, "libpn2_b200%s.so" % SUFFIX)
OBJ_DIR = os.path.join(HERE, "csrc", "build" + SUFFIX)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
] + os.environ.get("PN2_BUILD_DEFINES", "").split()


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps_mtime():
    files = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h")) + [os.path.abspath(__file__)]
    return max(os.path.getmtime(f) for f in files)


def build(verbose=False, force=False):
    """Compile every csrc/*.cu and link libpn2_b200.so.  Returns the library path."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_mtime = max([os.path.getmtime(f) for f in glob.glob(os.path.join(CSRC, "*.cuh")) +
                     glob.glob(os.path.join(HERE, "..", "include", "*.h"))] + [os.path.getmtime(__file__)])
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_mtime):
            continue
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s\n" % src)
    if failed:
        raise RuntimeError("libpn2_b200 build failed")
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lcuda"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("libpn2_b200 link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))
