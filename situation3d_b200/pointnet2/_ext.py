"""``pointnet2._ext`` -- the nine operators the reference exposes through pybind11
(lib/pointnet2/_ext_src/src/bindings.cpp:6-19), re-created over libpn2_b200's C ABI.

Same names, positional arguments, dtypes, layouts and failure behaviour as the
reference's ATen wrappers: every tensor must be contiguous, float32 / int32 and on
a CUDA device (utils.h:5-25 -> RuntimeError), outputs are freshly allocated and
zero-filled (sampling.cpp:25-27 etc.), launches are asynchronous on the current
stream.  Two deliberate differences: launch failures raise instead of calling
exit(-1) (cuda_utils.h:30-39), and the launch happens on the tensors' device
rather than on whatever device is current.
"""
import torch

from .. import _lib
from .._lib import check, lib, ptr, stream_ptr


def _check_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (CPU not supported)" % name)


def _check_contiguous(t, name):
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)


def _check_float(t, name):
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be a float tensor" % name)


def _check_int(t, name):
    if t.dtype != torch.int32:
        raise RuntimeError("%s must be an int tensor" % name)


def _f(t, name):
    _check_contiguous(t, name)
    _check_float(t, name)
    _check_cuda(t, name)


def _i(t, name):
    _check_contiguous(t, name)
    _check_int(t, name)
    _check_cuda(t, name)


def gather_points(points, idx):
    """points (B,C,N) f32, idx (B,m) i32 -> (B,C,m).  sampling.cpp:15-38"""
    _f(points, "points")
    _i(idx, "idx")
    B, C, N = points.shape
    m = idx.size(1)
    out = torch.zeros((B, C, m), device=points.device, dtype=torch.float32)
    with torch.cuda.device(points.device):
        check(lib.pn2_gather_points(B, C, N, m, ptr(points), ptr(idx), ptr(out), stream_ptr()), "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    """grad_out (B,C,m), idx (B,m) -> (B,C,n).  sampling.cpp:40-64"""
    _f(grad_out, "grad_out")
    _i(idx, "idx")
    B, C, m = grad_out.shape
    out = torch.zeros((B, C, n), device=grad_out.device, dtype=torch.float32)
    with torch.cuda.device(grad_out.device):
        check(lib.pn2_gather_points_grad(B, C, n, m, ptr(grad_out), ptr(idx), ptr(out), stream_ptr()),
              "gather_points_grad")
    return out


def furthest_point_sampling(points, nsamples):
    """points (B,N,3) f32 -> (B,nsamples) i32.  sampling.cpp:66-87"""
    _f(points, "points")
    B, N = points.size(0), points.size(1)
    out = torch.zeros((B, nsamples), device=points.device, dtype=torch.int32)
    # scratch like the reference's `temp` (sampling.cpp:74-76): large scenes run the bucketed kernel in it
    nbytes = lib.pn2_furthest_point_sampling_workspace_bytes(B, N, nsamples)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=points.device) if nbytes else None
    with torch.cuda.device(points.device):
        check(lib.pn2_furthest_point_sampling(B, N, nsamples, ptr(points), ptr(ws), nbytes, ptr(out), stream_ptr()),
              "furthest_point_sampling")
    return out


def three_nn(unknowns, knows):
    """unknowns (B,n,3), knows (B,m,3) -> [dist2 (B,n,3) f32, idx (B,n,3) i32].  interpolate.cpp:14-40"""
    _f(unknowns, "unknowns")
    _f(knows, "knows")
    B, n = unknowns.size(0), unknowns.size(1)
    m = knows.size(1)
    idx = torch.zeros((B, n, 3), device=unknowns.device, dtype=torch.int32)
    dist2 = torch.zeros((B, n, 3), device=unknowns.device, dtype=torch.float32)
    with torch.cuda.device(unknowns.device):
        check(lib.pn2_three_nn(B, n, m, ptr(unknowns), ptr(knows), ptr(dist2), ptr(idx), stream_ptr()), "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """points (B,c,m), idx (B,n,3), weight (B,n,3) -> (B,c,n).  interpolate.cpp:42-69"""
    _f(points, "points")
    _i(idx, "idx")
    _f(weight, "weight")
    B, c, m = points.shape
    n = idx.size(1)
    out = torch.zeros((B, c, n), device=points.device, dtype=torch.float32)
    with torch.cuda.device(points.device):
        check(lib.pn2_three_interpolate(B, c, m, n, ptr(points), ptr(idx), ptr(weight), ptr(out), stream_ptr()),
              "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """grad_out (B,c,n), idx, weight -> (B,c,m).  interpolate.cpp:71-99"""
    _f(grad_out, "grad_out")
    _i(idx, "idx")
    _f(weight, "weight")
    B, c, n = grad_out.shape
    out = torch.zeros((B, c, m), device=grad_out.device, dtype=torch.float32)
    with torch.cuda.device(grad_out.device):
        check(lib.pn2_three_interpolate_grad(B, c, n, m, ptr(grad_out), ptr(idx), ptr(weight), ptr(out),
                                             stream_ptr()), "three_interpolate_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz (B,m,3) centres FIRST, xyz (B,N,3) -> (B,m,nsample) i32.  ball_query.cpp:8-32"""
    _f(new_xyz, "new_xyz")
    _f(xyz, "xyz")
    B, m = new_xyz.size(0), new_xyz.size(1)
    N = xyz.size(1)
    idx = torch.zeros((B, m, nsample), device=new_xyz.device, dtype=torch.int32)
    from ..fused import ball_query as _ball_query      # workspace handling lives there
    return _ball_query(xyz, new_xyz, radius, nsample, out=idx)


def group_points(points, idx):
    """points (B,C,N), idx (B,npoints,nsample) -> (B,C,npoints,nsample).  group_points.cpp:12-36"""
    _f(points, "points")
    _i(idx, "idx")
    B, C, N = points.shape
    npoints, nsample = idx.size(1), idx.size(2)
    out = torch.zeros((B, C, npoints, nsample), device=points.device, dtype=torch.float32)
    with torch.cuda.device(points.device):
        check(lib.pn2_group_points(B, C, N, npoints, nsample, ptr(points), ptr(idx), ptr(out), stream_ptr()),
              "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    """grad_out (B,C,npoints,nsample), idx -> (B,C,n).  group_points.cpp:38-62"""
    _f(grad_out, "grad_out")
    _i(idx, "idx")
    B, C, npoints, nsample = grad_out.shape
    out = torch.zeros((B, C, n), device=grad_out.device, dtype=torch.float32)
    with torch.cuda.device(grad_out.device):
        check(lib.pn2_group_points_grad(B, C, n, npoints, nsample, ptr(grad_out), ptr(idx), ptr(out),
                                        stream_ptr()), "group_points_grad")
    return out


__all__ = ["gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
           "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"]
