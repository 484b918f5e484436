"""Training-mode forward of the backbone in the channel-last row layout (BASELINE.json config 4).

The reference trains through ``QueryAndGroup`` -> ``(B, C+3, npoint, nsample)`` NCHW tensor -> 1x1 ``Conv2d`` ->
``BatchNorm2d`` -> ReLU -> ``max_pool2d`` (pointnet2_utils.py:348-359, pytorch_utils.py:11-121,
pointnet2_modules.py:251-262): 553 MB per SA1 activation at B = 8, strided channel-first gathers and float atomics on
the way back (group_points_gpu.cu:43-64).  The same mathematics on rows -- one (row = centre x sample, C+3) matrix per
layer -- is a row gather (backward: row scatter-add), a GEMM per 1x1 convolution, BatchNorm over the rows (identical
statistics: BatchNorm2d normalises over B x npoint x nsample, which is exactly the row axis) and a max over each
centre's ``nsample`` consecutive rows.  Indices come from this library's sampling / ball-query / three-NN kernels
(non-differentiable, like the reference's); the dense algebra and its autograd are PyTorch's (cuBLAS), as the
reference's are cuDNN's.  Parameters and buffers are the modules' own (same names, same state_dict), BatchNorm
running statistics are updated exactly as ``nn.BatchNorm2d`` would.
"""
import ctypes
import os
import weakref

import torch
import torch.nn.functional as F

from . import fused as _fused
from ._lib import check, lib, ptr, stream_ptr


def rows_supported(mlp):
    """True when every layer is conv(1x1) [+ BatchNorm2d] + ReLU in that order (what SharedMLP builds by default)."""
    for layer in mlp:
        names = [k for k, _ in layer.named_children()]
        if not names or names[0] != "conv" or any(k not in ("conv", "bn", "activation") for k in names):
            return False
        conv = layer.conv
        if tuple(conv.kernel_size) != (1, 1) or tuple(conv.stride) != (1, 1) or tuple(conv.padding) != (0, 0):
            return False
        if "activation" in names and not isinstance(layer.activation, torch.nn.ReLU):
            return False
    return True


class _OnDevice:
    """``with _OnDevice(dev):`` -- torch.cuda.device(dev) only when ``dev`` is not already current (the training step is
    host-bound: ~350 launches per step, and the guard's two cudaSetDevice calls per launch add up)."""
    __slots__ = ("ctx",)

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


class _BnReluRows(torch.autograd.Function):
    """Training-mode BatchNorm (batch statistics, biased variance) + ReLU [+ max over each group of ``pool_ns`` consecutive
    rows] on a (rows, C) matrix: csrc/train_rows.cu.  ``running_mean`` / ``running_var`` (or None) are updated in place
    as nn.BatchNorm2d does.  Nothing but ``x`` is kept for the backward pass (the ReLU mask is recomputed from it)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, pool_ns, momentum, running_mean, running_var):
        R, C = x.shape
        dev = x.device
        st = stream_ptr()
        partials = torch.empty(lib.pn2_rows_bn_partials_bytes(R, C) // 8, dtype=torch.float64, device=dev)
        coef = torch.empty((2, C), dtype=torch.float32, device=dev)          # a | b
        stat = torch.empty((2, C), dtype=torch.float64, device=dev)          # mean | invstd
        nparts = ctypes.c_int(0)
        w, b_ = weight.detach().contiguous(), bias.detach().contiguous()
        pa = ctypes.c_void_p(coef.data_ptr())
        pb = ctypes.c_void_p(coef.data_ptr() + 4 * C)
        with _OnDevice(dev):
            check(lib.pn2_rows_bn_stats(R, C, ptr(x), ptr(partials), ctypes.byref(nparts), st), "rows_bn_stats")
            check(lib.pn2_rows_bn_finalize(C, nparts.value, ptr(partials), R, float(eps), ptr(w), ptr(b_), float(momentum),
                                           ptr(running_mean), ptr(running_var), pa, pb, ptr(stat),
                                           st), "rows_bn_finalize")
            if pool_ns:
                groups = R // pool_ns
                out = torch.empty((groups, C), dtype=torch.float32, device=dev)
                arg = torch.empty((groups, C), dtype=torch.uint8, device=dev)
                check(lib.pn2_rows_bn_relu_pool(groups, pool_ns, C, ptr(x), pa, pb, ptr(out), ptr(arg),
                                                st), "rows_bn_relu_pool")
            else:
                out = torch.empty_like(x)
                arg = None
                check(lib.pn2_rows_bn_relu_apply(R, C, ptr(x), pa, pb, ptr(out), st),
                      "rows_bn_relu_apply")
        ctx.pool_ns = pool_ns
        ctx.save_for_backward(x, w, coef, stat, arg)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, coef, stat, arg = ctx.saved_tensors
        R, C = x.shape
        dev = x.device
        st = stream_ptr()
        dout = dout.contiguous()
        partials = torch.empty(lib.pn2_rows_bn_partials_bytes(R, C) // 8, dtype=torch.float64, device=dev)
        k = torch.empty((5, C), dtype=torch.float32, device=dev)             # k1 | k2 | k3 | dgamma | dbeta
        dx = torch.empty_like(x)
        nparts = ctypes.c_int(0)
        ns = ctx.pool_ns
        a, b = ctypes.c_void_p(coef.data_ptr()), ctypes.c_void_p(coef.data_ptr() + 4 * C)
        kp = [ctypes.c_void_p(k.data_ptr() + 4 * C * i) for i in range(5)]
        with _OnDevice(dev):
            if ns:
                check(lib.pn2_rows_bn_relu_pool_bwd_reduce(R // ns, ns, C, ptr(dout), ptr(x), ptr(arg), a, b, ptr(partials),
                                                           ctypes.byref(nparts), st), "rows_bn_relu_pool_bwd_reduce")
            else:
                check(lib.pn2_rows_bn_relu_bwd_reduce(R, C, ptr(dout), ptr(x), a, b, ptr(partials), ctypes.byref(nparts),
                                                      st), "rows_bn_relu_bwd_reduce")
            check(lib.pn2_rows_bn_bwd_finalize(C, nparts.value, ptr(partials), R, ptr(stat), ptr(w), kp[0], kp[1],
                                               kp[2], kp[3], kp[4], st), "rows_bn_bwd_finalize")
            if ns:
                check(lib.pn2_rows_bn_relu_pool_bwd_apply(R // ns, ns, C, ptr(dout), ptr(x), ptr(arg), a, b, kp[0], kp[1],
                                                          kp[2], ptr(dx), st), "rows_bn_relu_pool_bwd_apply")
            else:
                check(lib.pn2_rows_bn_relu_bwd_apply(R, C, ptr(dout), ptr(x), a, b, kp[0], kp[1], kp[2], ptr(dx),
                                                     st), "rows_bn_relu_bwd_apply")
        return dx, k[3], k[4], None, None, None, None, None


class _GroupRows(torch.autograd.Function):
    """QueryAndGroup on rows (pointnet2_utils.py:348-359) in one pass: (rows, C+3) = [normalised relative xyz | features of
    the neighbours].  Backward: row scatter-add of the feature columns (the coordinates carry no gradient)."""

    @staticmethod
    def forward(ctx, rows, xyz, new_xyz, idx, radius, normalize):
        B, N, ld = rows.shape[0], rows.shape[1], rows.stride(1)
        C = rows.shape[2]
        npoint, ns = idx.shape[1], idx.shape[2]
        out = torch.empty((B * npoint * ns, C + 3), dtype=torch.float32, device=rows.device)
        st = stream_ptr()
        with _OnDevice(rows.device):
            check(lib.pn2_group_rows(B, N, npoint, ns, C, ld, ptr(idx), ptr(xyz), ptr(new_xyz), ptr(rows), float(radius),
                                     int(bool(normalize)), ptr(out), st), "group_rows")
        ctx.save_for_backward(idx)
        ctx.shape = (B, N, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None, None
        (idx,) = ctx.saved_tensors
        B, N, C = ctx.shape
        flat = (idx.long() + (torch.arange(B, device=idx.device) * N)[:, None, None]).reshape(-1)
        d = torch.zeros((B * N, C), dtype=dout.dtype, device=dout.device)
        d.index_add_(0, flat, dout[:, 3:])
        return d.view(B, N, C), None, None, None, None, None


def _group_rows_ok(rows, xyz):
    """rows (B, N, C) fp32 CUDA whose rows are contiguous (a column slice of point_clouds qualifies: pitch = stride(1))."""
    return (rows is not None and rows.is_cuda and rows.dtype == torch.float32 and rows.dim() == 3 and rows.stride(2) == 1 and
            rows.stride(0) == rows.shape[1] * rows.stride(1) and os.environ.get("PN2_TRAIN_GROUP_KERNEL", "1") != "0")


def _fused_bn_relu_ok(layer, x, pool_ns):
    """The hand-written kernels cover what SharedMLP builds: affine BatchNorm2d in training mode followed by ReLU, fp32
    CUDA rows, channel counts the kernels' thread layout divides."""
    if not (hasattr(layer, "bn") and hasattr(layer, "activation")) or os.environ.get("PN2_TRAIN_FUSED_BN", "1") == "0":
        return False
    bn = layer.bn.bn
    if not bn.training or bn.weight is None or bn.bias is None:
        return False
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2:
        return False
    R, C = x.shape
    if pool_ns and R % pool_ns:
        return False
    return bool(lib.pn2_rows_bn_supported(R, C, int(pool_ns)))


def shared_mlp_rows(mlp, x, pool_ns=0, first_linear_done=False):
    """SharedMLP (pytorch_utils.py:11-36) applied to (rows, Cin) instead of (B, Cin, npoint, nsample).  With ``pool_ns``
    the max over each group of ``pool_ns`` consecutive rows (pointnet2_modules.py:259-262) is taken after the last layer
    and (groups, Cout) is returned.  ``first_linear_done``: x already is the first layer's convolution output."""
    layers = list(mlp)
    for li, layer in enumerate(layers):
        conv = layer.conv
        if not (first_linear_done and li == 0):
            x = F.linear(x, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias)
        pool_here = pool_ns if li == len(layers) - 1 else 0
        if _fused_bn_relu_ok(layer, x, pool_here):
            bn = layer.bn.bn
            track = bn.track_running_stats and bn.running_mean is not None
            f = 0.0
            if track:
                bn.num_batches_tracked.add_(1)
                f = 1.0 / float(bn.num_batches_tracked) if bn.momentum is None else bn.momentum
            x = _BnReluRows.apply(x.contiguous(), bn.weight, bn.bias, float(bn.eps), int(pool_here), f,
                                  bn.running_mean if track else None, bn.running_var if track else None)
            continue
        if hasattr(layer, "bn") and layer.bn.bn.training and os.environ.get("PN2_TRAIN_FUSED_BN") == "2":
            # diagnostic (scripts/train_parity_diag.py): the same BatchNorm written as x*a + b in plain torch ops --
            # equivalent arithmetic, different rounding -- to tell rounding sensitivity from a kernel defect
            bn = layer.bn.bn
            mean, var = x.mean(dim=0), x.var(dim=0, unbiased=False)
            a_ = bn.weight * torch.rsqrt(var + bn.eps)
            x = x * a_ + (bn.bias - mean * a_)
        elif hasattr(layer, "bn"):
            bn = layer.bn.bn
            factor = 0.0
            if bn.training and bn.track_running_stats:
                bn.num_batches_tracked.add_(1)
                factor = 1.0 / float(bn.num_batches_tracked) if bn.momentum is None else bn.momentum
            x = F.batch_norm(x, bn.running_mean if (not bn.training or bn.track_running_stats) else None,
                             bn.running_var if (not bn.training or bn.track_running_stats) else None,
                             bn.weight, bn.bias, bn.training or bn.running_mean is None, factor, bn.eps)
        if hasattr(layer, "activation"):
            x = F.relu(x)
        if pool_here:
            x = x.view(x.shape[0] // pool_here, pool_here, x.shape[1]).amax(dim=1)
    return x


_SIDE_STREAMS = {}
_PREFETCHED = weakref.WeakKeyDictionary()      # net -> SamplingPyramid started by Pointnet2Backbone.prefetch_sampling


class SamplingPyramid:
    """The FPS chain of the four SA levels for one batch, enqueued on a side stream: it needs the coordinates only,
    so SA1's dense algebra overlaps the sampling of SA2-SA4, and ``Pointnet2Backbone.prefetch_sampling(next_batch)``
    lets the whole chain of the NEXT batch run under this batch's backward pass.  ``ready[l]`` is recorded when level
    l's indices and centres are complete; consumers wait for it on their own stream."""

    def __init__(self, net, pc):
        from .streams import sampling_stream
        mods = (net.sa1, net.sa2, net.sa3, net.sa4)
        B, N, W = pc.shape
        dev = pc.device
        self.source = pc
        self.xyz = torch.empty((B, N, 3), dtype=torch.float32, device=dev)      # filled by the first sampling kernel
        self.inds = [torch.empty((B, m.npoint), dtype=torch.int32, device=dev) for m in mods]
        self.cxyz = [torch.empty((B, m.npoint, 3), dtype=torch.float32, device=dev) for m in mods]
        self.ready = [torch.cuda.Event() for _ in mods]
        main = torch.cuda.current_stream(dev)
        side = _SIDE_STREAMS.get(dev.index)        # one per device; kept off the module (streams cannot be deep-copied)
        if side is None:
            side = _SIDE_STREAMS[dev.index] = sampling_stream(dev)
        side.wait_stream(main)                     # the batch (and the buffers above) are ready on the caller's stream
        with torch.cuda.stream(side), torch.no_grad():
            src = None
            for lvl in range(len(mods)):
                if lvl == 0:
                    _fused.fps_rows_into(pc, self.inds[0], self.cxyz[0], self.xyz)
                else:
                    _fused.fps_into(src, self.inds[lvl], self.cxyz[lvl])
                self.ready[lvl].record(side)
                src = self.cxyz[lvl]
        for t in [pc, self.xyz] + self.inds + self.cxyz:
            t.record_stream(side)

    def matches(self, pc):
        return pc is self.source or (pc.data_ptr() == self.source.data_ptr() and pc.shape == self.source.shape and
                                     pc._version == self.source._version)


def sa_rows(m, xyz, rows, sampled=None):
    """PointnetSAModuleVotes.forward (pointnet2_modules.py:210-277, max pooling) on rows (B, N, C).
    ``sampled`` = (inds, new_xyz) already computed (SamplingPyramid).
    Returns (new_xyz (B,np,3), out_rows (B,np,Cout), inds (B,np) i32)."""
    B, N, _ = xyz.shape
    with torch.no_grad():
        inds, new_xyz = sampled if sampled is not None else _fused.fps_with_xyz(xyz, m.npoint)
        idx = _fused.ball_query(xyz, new_xyz, m.radius, m.nsample)
        flat = (idx.long() + (torch.arange(B, device=xyz.device) * N)[:, None, None]).reshape(-1)
        gxyz = xyz.reshape(B * N, 3).index_select(0, flat).view(B, m.npoint, m.nsample, 3) - new_xyz[:, :, None, :]
        if m.normalize_xyz:
            gxyz = gxyz / m.radius                                                   # pointnet2_utils.py:350-351
        gxyz = gxyz.reshape(-1, 3)
    conv0 = m.mlp_module[0].conv
    if rows is not None and os.environ.get("PN2_TRAIN_SPLIT_L1", "0") == "1":
        # Layer 1 is linear in [xyz | features[idx]]: the feature part is applied ONCE PER POINT (B*N rows instead of
        # B*npoint*nsample: 3-16x fewer rows) and the (narrower) products are gathered; the three coordinate columns are
        # added per sample.  Same sum, different association (as the forward-only path's pn2_lin_tc_forward split);
        # the (rows, C+3) grouped matrix of the reference (pointnet2_utils.py:355-360) is never built.
        W = conv0.weight.view(conv0.out_channels, conv0.in_channels)
        Wf = W[:, 3:] if m.use_xyz else W
        P = F.linear(rows.reshape(B * N, rows.shape[2]), Wf)                         # (B*N, C1)
        z = P.index_select(0, flat)                                                  # backward: row scatter-add into dP
        if m.use_xyz:
            z = z.addmm_(gxyz, W[:, :3].t())
        if conv0.bias is not None:
            z = z + conv0.bias
        x = shared_mlp_rows(m.mlp_module, z, pool_ns=m.nsample, first_linear_done=True)
    elif m.use_xyz and _group_rows_ok(rows, xyz):
        x = _GroupRows.apply(rows, xyz, new_xyz, idx, m.radius, m.normalize_xyz)
        x = shared_mlp_rows(m.mlp_module, x, pool_ns=m.nsample)
    else:
        if rows is not None:
            g = rows.reshape(B * N, rows.shape[2]).index_select(0, flat)             # backward: row scatter-add
            x = torch.cat([gxyz, g], dim=1) if m.use_xyz else g                      # xyz channels first (:355-360)
        else:
            x = gxyz
        x = shared_mlp_rows(m.mlp_module, x, pool_ns=m.nsample)
    out_rows = x.view(B, m.npoint, x.shape[1])
    return new_xyz, out_rows, inds


def fp_rows(m, unknown, known, skip_rows, known_rows):
    """PointnetFPModule.forward (pointnet2_modules.py:376-421) on rows: skip_rows (B,n,C1) or None, known_rows (B,m,C2)."""
    B, n, _ = unknown.shape
    mk = known.shape[1]
    with torch.no_grad():
        d2, i3 = _fused.three_nn(unknown, known)
        recip = 1.0 / (torch.sqrt(d2) + 1e-8)                                        # pointnet2_utils.py:142, :400
        w = recip / recip.sum(dim=2, keepdim=True)
        flat = (i3.long() + (torch.arange(B, device=unknown.device) * mk)[:, None, None]).reshape(-1)
    g = known_rows.reshape(B * mk, known_rows.shape[2]).index_select(0, flat).view(B * n, 3, known_rows.shape[2])
    x = (g * w.reshape(B * n, 3, 1)).sum(dim=1)
    if skip_rows is not None:
        x = torch.cat([x, skip_rows.reshape(B * n, skip_rows.shape[2])], dim=1)      # interpolated channels first (:412-416)
    x = shared_mlp_rows(m.mlp, x)
    return x.view(B, n, x.shape[1])


def backbone_forward_rows(net, pc, data_dict):
    """Pointnet2Backbone.forward in training mode over rows; same dictionary keys as the other paths."""
    pyr = _PREFETCHED.pop(net, None)
    if pyr is None or not pyr.matches(pc):
        pyr = SamplingPyramid(net, pc.contiguous())
    main = torch.cuda.current_stream(pc.device)
    main.wait_event(pyr.ready[0])
    xyz = pyr.xyz
    rows = pc[..., 3:] if pc.shape[2] > 3 else None
    level_xyz, level_rows = [], []
    for lvl, m in enumerate((net.sa1, net.sa2, net.sa3, net.sa4), start=1):
        main.wait_event(pyr.ready[lvl - 1])
        xyz, rows, inds = sa_rows(m, xyz, rows, sampled=(pyr.inds[lvl - 1], pyr.cxyz[lvl - 1]))
        level_xyz.append(xyz)
        level_rows.append(rows)
        data_dict["sa%d_inds" % lvl] = inds
        data_dict["sa%d_xyz" % lvl] = xyz
        data_dict["sa%d_features" % lvl] = rows.transpose(1, 2)
    f = fp_rows(net.fp1, level_xyz[2], level_xyz[3], level_rows[2], level_rows[3])
    f = fp_rows(net.fp2, level_xyz[1], level_xyz[2], level_rows[1], f)
    data_dict["fp2_features"] = f.transpose(1, 2)
    data_dict["fp2_xyz"] = level_xyz[1]
    data_dict["fp2_inds"] = data_dict["sa1_inds"][:, 0:level_xyz[1].shape[1]]
    return data_dict
