"""Host-side plumbing of the fused CUDA layers: BatchNorm folding, weight images, launches.

Nothing here computes on the host: it reads module parameters, folds eval-mode BatchNorm
into the 1x1-conv weights (W' = W*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps)),
hands them to the C ABI's pack functions and launches the fused kernels on the current stream.
"""
import ctypes

import contextlib

import torch
import torch.nn as nn

from ._lib import Pn2Error, check, int_array, lib, ptr, ptr_array, stream_ptr


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
        self.f32_only = True
        self._f32_alt = None
        self._split = {}

    def get(self, mlp, precision, kind="sa"):
        key = (precision, kind) + _mlp_state_key(mlp)
        if key != self.key:
            folded = fold_shared_mlp(mlp)
            if folded is None:
                self.key, self.image, self.dims, self.folded = key, None, None, None
                return None
            dims = [folded[0][0].shape[1]] + [w.shape[0] for w, _ in folded]
            packer = _PACKERS[precision] if kind == "sa" else (_pack_fp_bf16 if precision == "bf16" else _pack_f32)
            self.image = packer(dims, folded)
            if self.image is None and precision == "bf16":
                self.image = _pack_f32(dims, folded)      # shapes outside the tcgen05 kernels: fp32 kernel
                self.f32_only = True
            else:
                self.f32_only = precision != "bf16"
            self.key, self.dims, self.folded = key, dims, folded
            self._f32_alt = None
            self._split = {}
        return self

    def split(self, kin, skip, x_is_bf16):
        """Images of the split first layer (csrc/lin_tc.cu): (lin image, SA image over the c1-wide P rows), or
        None when the shapes are outside the kernels.  Cached per input layout."""
        key = (kin, skip, x_is_bf16)
        if key not in self._split:
            self._split[key] = _pack_split(self.dims, self.folded, kin, skip, x_is_bf16)
        return self._split[key]


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


def _pack_bf16(dims, folded):
    """Weight image of the tcgen05 SA kernel (3 layers, xyz columns first in layer 1); None when the
    widths are outside what that kernel covers (such stacks run on the fp32 kernel)."""
    if len(folded) != 3:
        return None
    c, c1, c2, c3 = dims[0] - 3, dims[1], dims[2], dims[3]
    nbytes = lib.pn2_sa_tc_weight_image_bytes(c, c1, c2, c3)
    if nbytes == 0:
        return None
    dev = folded[0][0].device
    image = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    (w1, b1), (w2, b2), (w3, b3) = folded
    with torch.cuda.device(dev):
        check(lib.pn2_sa_tc_pack_weights(c, c1, c2, c3, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(b3),
                                         ptr(image), stream_ptr()), "sa_tc_pack_weights")
    return image


def _pack_split(dims, folded, kin, skip, x_is_bf16):
    """W1 [xyz ; f] = W1[:, :3] xyz + W1[:, 3:] f: the feature half becomes a per-point row GEMM (lin image),
    the SA kernel keeps the xyz / bias columns and an identity block over the c1-wide result."""
    if len(folded) != 3:
        return None
    c, c1, c2, c3 = dims[0] - 3, dims[1], dims[2], dims[3]
    if not lib.pn2_lin_tc_supported(kin, c1, 1 if x_is_bf16 else 0):
        return None
    sa_bytes = lib.pn2_sa_tc_weight_image_bytes(c1, c1, c2, c3)
    lin_bytes = lib.pn2_lin_tc_weight_image_bytes(kin, c1)
    if sa_bytes == 0 or lin_bytes == 0:
        return None
    (w1, b1), (w2, b2), (w3, b3) = folded
    dev = w1.device
    lin_image = torch.empty(lin_bytes, dtype=torch.uint8, device=dev)
    sa_image = torch.empty(sa_bytes, dtype=torch.uint8, device=dev)
    w1p = torch.cat([w1[:, :3], torch.eye(c1, dtype=torch.float32, device=dev)], dim=1).contiguous()
    with torch.cuda.device(dev):
        check(lib.pn2_lin_tc_pack_weights(kin, skip, c, c1, ptr(w1), c + 3, 3, ptr(lin_image), stream_ptr()),
              "lin_tc_pack_weights")
        check(lib.pn2_sa_tc_pack_weights(c1, c1, c2, c3, ptr(w1p), ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(b3),
                                         ptr(sa_image), stream_ptr()), "sa_tc_pack_weights")
    return lin_image, sa_image


def _rup(v, m):
    return (v + m - 1) // m * m


def split_first_layer(c, c1):
    """True when gathering c1-wide per-point results is cheaper than gathering the c-wide feature rows
    (layer-1 K of the fused kernel: c1 + 16 against round_up(round_up(c, 8) + 8, 16))."""
    import os
    if os.environ.get("PN2_SA_SPLIT", "1") == "0":
        return False
    return c1 in (64, 128) and c1 + 16 < _rup(_rup(c, 8) + 8, 16)


def lin_rows(lin_image, x, base_ptr, rows, kin, c1, x_is_bf16, ld):
    """P (rows, c1) bf16 = X W^T over channel-last rows (pn2_lin_tc_forward)."""
    out = torch.empty((rows, c1), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.pn2_lin_tc_forward(rows, kin, c1, base_ptr, 1 if x_is_bf16 else 0, ld, ptr(lin_image), ptr(out),
                                     stream_ptr()), "lin_tc_forward")
    return out


def _pack_fp_bf16(dims, folded):
    """Weight image of the tcgen05 FP kernel (2 layers); None when the widths are outside its coverage."""
    if len(folded) != 2:
        return None
    k0, c1, c2 = dims
    nbytes = lib.pn2_fp_tc_weight_image_bytes(k0, 0, c1, c2)       # the image depends on c_known + c_skip only
    if nbytes == 0:
        return None
    dev = folded[0][0].device
    image = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    (w1, b1), (w2, b2) = folded
    with torch.cuda.device(dev):
        check(lib.pn2_fp_tc_pack_weights(k0, 0, c1, c2, ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(image), stream_ptr()),
              "fp_tc_pack_weights")
    return image


_PACKERS = {"fp32": _pack_f32, "bf16": _pack_bf16}


def rows_from_channels(features):
    """(B,C,N) -> (B,N,C) channel-last rows."""
    B, C, N = features.shape
    rows = torch.empty((B, N, C), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        check(lib.pn2_rows_from_channels(B, C, N, ptr(features), ptr(rows), stream_ptr()), "rows_from_channels")
    return rows


_SAMPLING_MODE = [0]          # PN2_FPS_LATENCY; see sampling_mode()


@contextlib.contextmanager
def sampling_mode(mode):
    """``with fused.sampling_mode("throughput"):`` -- sampling launches issued inside (and graphs captured inside) choose
    the kernel with the least SM time per scene instead of the fastest single launch (include/pn2_b200.h:
    PN2_FPS_THROUGHPUT; at 40 000 points the bucketed one-CTA-per-scene kernel instead of the 8-CTA cluster kernel).
    For callers that keep >= ~150 scenes in flight (graphs.BackbonePipeline decides by lanes x batch).  Results are
    identical in both modes."""
    value = {"latency": 0, "throughput": 1}[mode]
    prev = _SAMPLING_MODE[0]
    _SAMPLING_MODE[0] = value
    try:
        yield
    finally:
        _SAMPLING_MODE[0] = prev


def _fps_workspace(B, N, npoint, device):
    """Scratch of the bucketed sampling kernel (csrc/fps_bucket.cu), from PyTorch's caching allocator; None when the
    register-resident kernels run (small scenes; mid-sized scenes in latency mode)."""
    nbytes = lib.pn2_furthest_point_sampling_workspace_bytes_mode(B, N, npoint, _SAMPLING_MODE[0])
    return (torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes) if nbytes else (None, 0)


def fps_with_xyz(xyz, npoint):
    """Furthest point sampling that also returns new_xyz = xyz[inds] (B,npoint,3)."""
    B, N, _ = xyz.shape
    inds = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device)
    fps_into(xyz, inds, new_xyz)
    return inds, new_xyz


def fps_into(xyz, inds, new_xyz):
    """Same, into caller-allocated outputs (used by the backbone's side-stream sampling pyramid)."""
    B, N, _ = xyz.shape
    ws, nbytes = _fps_workspace(B, N, inds.shape[1], xyz.device)
    with torch.cuda.device(xyz.device):
        check(lib.pn2_furthest_point_sampling_xyz_ws(B, N, inds.shape[1], ptr(xyz), ptr(inds), ptr(new_xyz), ptr(ws),
                                                     nbytes, stream_ptr()), "furthest_point_sampling_xyz")


def fps_rows_into(rows, inds, new_xyz, xyz_copy):
    """Sampling straight from (B, N, pitch) rows whose first three floats are xyz (``point_clouds`` in place); also
    fills ``xyz_copy`` (B, N, 3), the contiguous coordinates the later kernels read."""
    B, N, pitch = rows.shape
    ws, nbytes = _fps_workspace(B, N, inds.shape[1], rows.device)
    with torch.cuda.device(rows.device):
        check(lib.pn2_furthest_point_sampling_rows_ws(B, N, inds.shape[1], ptr(rows), pitch, ptr(inds), ptr(new_xyz),
                                                      ptr(xyz_copy), ptr(ws), nbytes, stream_ptr()),
              "furthest_point_sampling_rows")


def ball_query(xyz, new_xyz, radius, nsample, out=None):
    """Ball query (uniform grid for large scenes, plain scan otherwise); the workspace comes from
    PyTorch's caching allocator."""
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = out if out is not None else torch.empty((B, m, nsample), dtype=torch.int32, device=xyz.device)
    nbytes = lib.pn2_ball_query_workspace_bytes(B, N, m, int(nsample))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=xyz.device) if nbytes else None
    with torch.cuda.device(xyz.device):
        check(lib.pn2_ball_query_ws(B, N, m, float(radius), int(nsample), ptr(new_xyz), ptr(xyz), ptr(idx),
                                    ptr(ws), nbytes, stream_ptr()), "ball_query")
    return idx


def ball_query_grid_build(points, n, m, radius, nsample):
    """Builds the uniform grid over `points` ((B, n, pitch) f32, xyz in the first three columns) on the
    current stream; returns the workspace for ball_query_grid_query, or None when this shape takes the
    plain scan.  The grid needs no centres, so it can be built while the sampling kernel still runs."""
    B, pitch = points.shape[0], points.shape[2]
    nbytes = lib.pn2_ball_query_workspace_bytes(B, n, m, int(nsample))
    if not nbytes or not radius > 0:
        return None
    ws = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
    with torch.cuda.device(points.device):
        check(lib.pn2_ball_query_grid_build(B, n, m, float(radius), int(nsample), ptr(points), pitch, ptr(ws),
                                            nbytes, stream_ptr()), "ball_query_grid_build")
    return ws


def ball_query_grid_query(ws, n, new_xyz, radius, nsample):
    B, m, _ = new_xyz.shape
    idx = torch.empty((B, m, nsample), dtype=torch.int32, device=new_xyz.device)
    with torch.cuda.device(new_xyz.device):
        check(lib.pn2_ball_query_grid_query(B, n, m, float(radius), int(nsample), ptr(new_xyz), ptr(idx), ptr(ws),
                                            ws.numel(), stream_ptr()), "ball_query_grid_query")
    return idx


def three_nn(unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
    with torch.cuda.device(unknown.device):
        check(lib.pn2_three_nn(B, n, m, ptr(unknown), ptr(known), ptr(dist2), ptr(idx), stream_ptr()), "three_nn")
    return dist2, idx


def sa_forward_f32(img, xyz, new_xyz, idx, table, ld, c, use_xyz, inv_radius, want_rows=True, raw_skip=0):
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


def bf16_rows(table, ld, c):
    """bf16 channel-last row table (B, n, row_elems) from f32 channel-last rows (pitch ld floats)."""
    B, n = table.shape[0], table.shape[1]
    row_elems = lib.pn2_sa_tc_row_elems(c)
    out = torch.empty((B, n, row_elems), dtype=torch.bfloat16, device=table.device)
    with torch.cuda.device(table.device):
        check(lib.pn2_sa_tc_pack_rows(B, n, c, ptr(table), n * ld, ld, ptr(out), stream_ptr()), "sa_tc_pack_rows")
    return out


def _f32_rows(rows):
    return rows if rows is None or rows.dtype == torch.float32 else rows.float()


def sa_bf16_table(img, table, ld, c, B, N, npoint, nsample, raw_skip=0):
    """What the tensor-core SA kernel gathers from, and the weight image that goes with it: the c1-wide
    per-point layer-1 rows (split first layer, csrc/lin_tc.cu) when that is the cheaper gather, else the bf16
    feature rows.  Returns (table (B,N,row_elems) bf16, weight image, channels per row)."""
    dims = img.dims
    c1 = dims[1]
    if split_first_layer(c, c1) and lib.pn2_sa_tc_supported(c1, c1, dims[2], dims[3], npoint, nsample):
        pair = None
        if table.dtype != torch.bfloat16:
            # f32 rows read in place: raw_skip elements precede feature 0 in each ld-pitched, 16-byte aligned row
            kin = _rup(raw_skip + c, 4)
            base = table.data_ptr() - 4 * raw_skip
            if ld % 4 == 0 and kin <= ld and base % 16 == 0 and table.stride() == (N * ld, ld, 1):
                pair = img.split(kin, raw_skip, False)
                if pair is not None:
                    table = lin_rows(pair[0], table, ctypes.c_void_p(base), B * N, kin, c1, False, ld).view(B, N, c1)
        if pair is None:
            bf16_skip = raw_skip if table.dtype == torch.bfloat16 else 0
            if table.dtype != torch.bfloat16:
                table = bf16_rows(table, ld, c)
            kin = table.shape[2]
            # bf16 rows may carry `raw_skip` unused leading elements (the compact transport format keeps every feature
            # in the column it has in the fp32 row, so both formats feed the tensor cores identical operands)
            pair = img.split(kin, bf16_skip, True)
            if pair is not None:
                table = lin_rows(pair[0], table, ptr(table), B * N, kin, c1, True, kin).view(B, N, c1)
        if pair is not None:
            return table, pair[1], c1
    if table.dtype == torch.bfloat16 and raw_skip:
        raise Pn2Error("bf16 feature rows with leading padding need the split first layer (csrc/lin_tc.cu)")
    if table.dtype != torch.bfloat16:
        table = bf16_rows(table, ld, c)
    return table, img.image, c


def sa_bf16_fused(dims, image, c_eff, xyz, new_xyz, idx, table, inv_radius, want_rows=True, out=None, out_rows=None):
    """pn2_sa_tc_forward over a prepared row table (sa_bf16_table)."""
    B, N, _ = xyz.shape
    npoint, nsample = idx.shape[1], idx.shape[2]
    cout = dims[3]
    if out is None:
        out = torch.empty((B, cout, npoint), dtype=torch.float32, device=xyz.device)
    if out_rows is None and want_rows:
        out_rows = torch.empty((B, npoint, cout), dtype=torch.bfloat16, device=xyz.device)
    with torch.cuda.device(xyz.device):
        check(lib.pn2_sa_tc_forward(B, N, npoint, nsample, c_eff, dims[1], dims[2], dims[3], float(inv_radius), ptr(xyz),
                                    ptr(new_xyz), ptr(table), ptr(idx), ptr(image), ptr(out), ptr(out_rows),
                                    stream_ptr()), "sa_tc_forward")
    return out, out_rows


def _f32_alternative(img):
    """fp32 image of a stack the tensor-core kernels do not cover for this call's shape; raises (instead of handing a
    NULL image to the C ABI) when the fp32 kernel does not cover it either."""
    alt = img if img.f32_only else img._f32_alt
    if alt is None:
        alt = MlpImage()
        alt.dims, alt.folded, alt.image = img.dims, img.folded, _pack_f32(img.dims, img.folded)
        img._f32_alt = alt
    if alt.image is None:
        raise Pn2Error("no fused kernel covers the SharedMLP %s with this layer shape; construct the module with "
                       "fused=False" % (img.dims,))
    return alt


def sa_forward_bf16(img, xyz, new_xyz, idx, table, ld, c, use_xyz, inv_radius, want_rows=True, raw_skip=0):
    """Fused SA layer on tcgen05 (pn2_sa_tc_forward).  ``table`` is either f32 channel-last rows (read in place by
    the per-point GEMM, or packed to bf16 here) or the bf16 row table a previous layer produced.  Shapes the
    tensor-core kernel does not cover run on the fp32 kernel (with an fp32 image built on demand)."""
    B, N, _ = xyz.shape
    npoint, nsample = idx.shape[1], idx.shape[2]
    dims = img.dims
    ok = use_xyz and not img.f32_only and len(dims) == 4 and \
        lib.pn2_sa_tc_supported(c, dims[1], dims[2], dims[3], npoint, nsample)
    if not ok:
        alt = _f32_alternative(img)
        rows = _f32_rows(table)
        if rows is not table:
            ld = rows.shape[2]
        return sa_forward_f32(alt, xyz, new_xyz, idx, rows, ld, c, use_xyz, inv_radius, want_rows)
    table, image, c_eff = sa_bf16_table(img, table, ld, c, B, N, npoint, nsample, raw_skip)
    return sa_bf16_fused(dims, image, c_eff, xyz, new_xyz, idx, table, inv_radius, want_rows)


def _bf16_rows(rows):
    return rows if rows is None or rows.dtype == torch.bfloat16 else rows.bfloat16()


def fp_forward_bf16(img, dist2, idx, known_rows, skip_rows, want_rows=True):
    """Fused FP layer on tcgen05 (pn2_fp_tc_forward) over bf16 channel-last rows; widths the tensor-core
    kernel does not cover run on the fp32 kernel."""
    B, n, _ = idx.shape
    m, c_known = known_rows.shape[1], known_rows.shape[2]
    c_skip = 0 if skip_rows is None else skip_rows.shape[2]
    dims = img.dims
    ok = not img.f32_only and len(dims) == 3 and lib.pn2_fp_tc_supported(c_known, c_skip, dims[1], dims[2])
    if not ok:
        alt = _f32_alternative(img)
        return fp_forward_f32(alt, dist2, idx, _f32_rows(known_rows), _f32_rows(skip_rows), want_rows)
    known_rows, skip_rows = _bf16_rows(known_rows).contiguous(), _bf16_rows(skip_rows)
    cout = dims[2]
    out = torch.empty((B, cout, n), dtype=torch.float32, device=idx.device)
    out_rows = torch.empty((B, n, cout), dtype=torch.bfloat16, device=idx.device) if want_rows else None
    with torch.cuda.device(idx.device):
        check(lib.pn2_fp_tc_forward(B, n, m, c_known, c_skip, dims[1], dims[2], ptr(dist2), ptr(idx), ptr(known_rows),
                                    ptr(skip_rows), ptr(img.image), ptr(out), ptr(out_rows), stream_ptr()),
              "fp_tc_forward")
    return out, out_rows


def fp_layer(precision, img, unknown, known, known_rows, skip_rows, want_rows=True):
    """One whole feature-propagation layer from coordinates: three_nn + interpolation + concat + SharedMLP.  The bf16
    arm runs it as ONE launch (pn2_fp_tc2_forward: 4-CTA clusters, resident weight quarters) when the shapes allow;
    otherwise three_nn followed by the fused MLP kernel of the precision."""
    import os
    B, n, _ = unknown.shape
    m = known.shape[1]
    dims = img.dims
    # The cluster kernel is the low-latency one (config 1: 20 us against 58 us for three_nn + fp_tc at B = 1) but holds four
    # SMs per 128-point tile; with many tiles in a launch (and batches overlapping on other streams) the single-CTA kernel
    # costs less SM time per layer (measured at B = 8, 8 batches in flight: 11.5 k against 10.1 k scenes/s).  PN2_FP_TC2
    # = 1 / 0 forces one or the other.
    mode = os.environ.get("PN2_FP_TC2", "auto")
    use_tc2 = mode == "1" or (mode == "auto" and B * ((n + 127) // 128) * 4 <= 64)
    if precision == "bf16" and not img.f32_only and len(dims) == 3 and use_tc2:
        c_known = known_rows.shape[2]
        c_skip = 0 if skip_rows is None else skip_rows.shape[2]
        if known_rows.dtype == torch.bfloat16 and (skip_rows is None or skip_rows.dtype == torch.bfloat16) and \
                lib.pn2_fp_tc2_supported(c_known, c_skip, dims[1], dims[2], m):
            known_rows = known_rows.contiguous()
            skip_rows = None if skip_rows is None else skip_rows.contiguous()
            out = torch.empty((B, dims[2], n), dtype=torch.float32, device=unknown.device)
            out_rows = torch.empty((B, n, dims[2]), dtype=torch.bfloat16, device=unknown.device) if want_rows else None
            with torch.cuda.device(unknown.device):
                check(lib.pn2_fp_tc2_forward(B, n, m, c_known, c_skip, dims[1], dims[2], ptr(unknown), ptr(known),
                                             ptr(known_rows), ptr(skip_rows), ptr(img.image), ptr(out), ptr(out_rows),
                                             stream_ptr()), "fp_tc2_forward")
            return out, out_rows
    dist2, idx = three_nn(unknown, known)
    return FP_FORWARD[precision](img, dist2, idx, known_rows, skip_rows, want_rows)


SA_FORWARD = {"fp32": sa_forward_f32, "bf16": sa_forward_bf16}
FP_FORWARD = {"fp32": fp_forward_f32, "bf16": fp_forward_bf16}
