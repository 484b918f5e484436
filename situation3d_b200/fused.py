"""Host-side plumbing of the fused CUDA layers: BatchNorm folding, weight images, launches.

Nothing here computes on the host: it reads module parameters, folds eval-mode BatchNorm
into the 1x1-conv weights (W' = W*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps)),
hands them to the C ABI's pack functions and launches the fused kernels on the current stream.
"""
import torch
import torch.nn as nn

from ._lib import check, int_array, lib, ptr, ptr_array, stream_ptr


def fold_shared_mlp(mlp):
    """SharedMLP (pytorch_utils.py:11-36) in eval mode -> [(W (cout,cin), b (cout))] fp32, or None
    when a layer is not conv(1x1) [+BN] + ReLU in that order (such stacks use the unfused path)."""
    layers = []
    for layer in mlp:
        mods = list(layer.named_children())
        names = [k for k, _ in mods]
        conv = dict(mods).get("conv")
        if conv is None or not isinstance(conv, nn.Conv2d) or names[0] != "conv":
            return None
        if tuple(conv.kernel_size) != (1, 1) or tuple(conv.stride) != (1, 1) or tuple(conv.padding) != (0, 0):
            return None
        act = dict(mods).get("activation")
        if not isinstance(act, nn.ReLU):
            return None
        w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).float()
        b = conv.bias.detach().float() if conv.bias is not None else torch.zeros_like(w[:, 0])
        bn_wrap = dict(mods).get("bn")
        if bn_wrap is not None:
            bn = bn_wrap[0]
            if bn.running_mean is None:
                return None
            scale = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
            if bn.weight is not None:
                scale = scale * bn.weight.detach().float()
            shift = bn.bias.detach().float() if bn.bias is not None else torch.zeros_like(scale)
            w = w * scale[:, None]
            b = (b - bn.running_mean.detach().float()) * scale + shift
        layers.append((w.contiguous(), b.contiguous()))
    return layers


def _mlp_state_key(mlp):
    return tuple((t.data_ptr(), t._version) for t in list(mlp.parameters()) + list(mlp.buffers()))


class MlpImage:
    """Device image of a folded SharedMLP for one precision, rebuilt when a parameter changes."""

    def __init__(self):
        self.key = None
        self.dims = None
        self.image = None
        self.folded = None

    def get(self, mlp, precision):
        key = (precision,) + _mlp_state_key(mlp)
        if key != self.key:
            folded = fold_shared_mlp(mlp)
            if folded is None:
                self.key, self.image, self.dims, self.folded = key, None, None, None
                return None
            dims = [folded[0][0].shape[1]] + [w.shape[0] for w, _ in folded]
            self.image = _PACKERS[precision](dims, folded)
            self.key, self.dims, self.folded = key, dims, folded
        return self


def _pack_f32(dims, folded):
    n = len(folded)
    cd = int_array(dims)
    if not lib.pn2_mlp_f32_supported(n, cd):
        return None
    nbytes = lib.pn2_mlp_f32_image_bytes(n, cd)
    dev = folded[0][0].device
    image = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    ws = [w for w, _ in folded]
    bs = [b for _, b in folded]
    with torch.cuda.device(dev):
        check(lib.pn2_mlp_f32_pack(n, cd, ptr_array(ws), ptr_array(bs), ptr(image), stream_ptr()), "mlp_f32_pack")
    return image


_PACKERS = {"fp32": _pack_f32}


def rows_from_channels(features):
    """(B,C,N) -> (B,N,C) channel-last rows."""
    B, C, N = features.shape
    rows = torch.empty((B, N, C), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        check(lib.pn2_rows_from_channels(B, C, N, ptr(features), ptr(rows), stream_ptr()), "rows_from_channels")
    return rows


def fps_with_xyz(xyz, npoint):
    """Furthest point sampling that also returns new_xyz = xyz[inds] (B,npoint,3)."""
    B, N, _ = xyz.shape
    inds = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        check(lib.pn2_furthest_point_sampling_xyz(B, N, npoint, ptr(xyz), ptr(inds), ptr(new_xyz), stream_ptr()),
              "furthest_point_sampling_xyz")
    return inds, new_xyz


def fps_into(xyz, inds, new_xyz):
    """Same, into caller-allocated outputs (used by the backbone's side-stream sampling pyramid)."""
    B, N, _ = xyz.shape
    with torch.cuda.device(xyz.device):
        check(lib.pn2_furthest_point_sampling_xyz(B, N, inds.shape[1], ptr(xyz), ptr(inds), ptr(new_xyz),
                                                  stream_ptr()), "furthest_point_sampling_xyz")


def ball_query(xyz, new_xyz, radius, nsample):
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.empty((B, m, nsample), dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        check(lib.pn2_ball_query(B, N, m, float(radius), int(nsample), ptr(new_xyz), ptr(xyz), ptr(idx),
                                 stream_ptr()), "ball_query")
    return idx


def three_nn(unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
    with torch.cuda.device(unknown.device):
        check(lib.pn2_three_nn(B, n, m, ptr(unknown), ptr(known), ptr(dist2), ptr(idx), stream_ptr()), "three_nn")
    return dist2, idx


def sa_forward_f32(img, xyz, new_xyz, idx, table, ld, c, use_xyz, inv_radius, want_rows=True):
    """Fused QueryAndGroup-gather + SharedMLP + max-pool (pn2_sa_forward_f32).
    table: tensor whose data pointer is the first feature of row (0,0); ld = row pitch in floats."""
    B, N, _ = xyz.shape
    npoint, nsample = idx.shape[1], idx.shape[2]
    cout = img.dims[-1]
    out = torch.empty((B, cout, npoint), dtype=torch.float32, device=xyz.device)
    out_rows = torch.empty((B, npoint, cout), dtype=torch.float32, device=xyz.device) if want_rows else None
    with torch.cuda.device(xyz.device):
        check(lib.pn2_sa_forward_f32(B, N, npoint, nsample, c, ptr(table), ld, 1 if use_xyz else 0,
                                     float(inv_radius), ptr(xyz), ptr(new_xyz), ptr(idx), len(img.dims) - 1,
                                     int_array(img.dims), ptr(img.image), ptr(out), ptr(out_rows), stream_ptr()),
              "sa_forward_f32")
    return out, out_rows


def fp_forward_f32(img, dist2, idx, known_rows, skip_rows, want_rows=True):
    """Fused 3-NN weights + three_interpolate + concat + SharedMLP (pn2_fp_forward_f32)."""
    B, n, _ = idx.shape
    m, c_known = known_rows.shape[1], known_rows.shape[2]
    c_skip = 0 if skip_rows is None else skip_rows.shape[2]
    cout = img.dims[-1]
    out = torch.empty((B, cout, n), dtype=torch.float32, device=idx.device)
    out_rows = torch.empty((B, n, cout), dtype=torch.float32, device=idx.device) if want_rows else None
    with torch.cuda.device(idx.device):
        check(lib.pn2_fp_forward_f32(B, n, m, c_known, c_skip, ptr(dist2), ptr(idx), ptr(known_rows),
                                     ptr(skip_rows), len(img.dims) - 1, int_array(img.dims), ptr(img.image),
                                     ptr(out), ptr(out_rows), stream_ptr()), "fp_forward_f32")
    return out, out_rows


SA_FORWARD = {"fp32": sa_forward_f32}
FP_FORWARD = {"fp32": fp_forward_f32}
