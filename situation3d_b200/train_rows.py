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
import torch
import torch.nn.functional as F

from . import fused as _fused


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


def shared_mlp_rows(mlp, x):
    """SharedMLP (pytorch_utils.py:11-36) applied to (rows, Cin) instead of (B, Cin, npoint, nsample)."""
    for layer in mlp:
        conv = layer.conv
        x = F.linear(x, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias)
        if hasattr(layer, "bn"):
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
    return x


def sa_rows(m, xyz, rows):
    """PointnetSAModuleVotes.forward (pointnet2_modules.py:210-277, max pooling) on rows (B, N, C).
    Returns (new_xyz (B,np,3), out_rows (B,np,Cout), inds (B,np) i32)."""
    B, N, _ = xyz.shape
    with torch.no_grad():
        inds, new_xyz = _fused.fps_with_xyz(xyz, m.npoint)
        idx = _fused.ball_query(xyz, new_xyz, m.radius, m.nsample)
        flat = (idx.long() + (torch.arange(B, device=xyz.device) * N)[:, None, None]).reshape(-1)
        gxyz = xyz.reshape(B * N, 3).index_select(0, flat).view(B, m.npoint, m.nsample, 3) - new_xyz[:, :, None, :]
        if m.normalize_xyz:
            gxyz = gxyz / m.radius                                                   # pointnet2_utils.py:350-351
        gxyz = gxyz.reshape(-1, 3)
    if rows is not None:
        g = rows.reshape(B * N, rows.shape[2]).index_select(0, flat)                 # backward: row scatter-add
        x = torch.cat([gxyz, g], dim=1) if m.use_xyz else g                          # xyz channels first (:355-360)
    else:
        x = gxyz
    x = shared_mlp_rows(m.mlp_module, x)
    out_rows = x.view(B * m.npoint, m.nsample, x.shape[1]).amax(dim=1).view(B, m.npoint, x.shape[1])
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
    xyz = pc[..., :3].contiguous()
    rows = pc[..., 3:] if pc.shape[2] > 3 else None
    level_xyz, level_rows = [], []
    for lvl, m in enumerate((net.sa1, net.sa2, net.sa3, net.sa4), start=1):
        xyz, rows, inds = sa_rows(m, xyz, rows)
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
