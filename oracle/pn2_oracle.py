"""Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file; nothing under situation3d_b200/ does.  It is the checker (and the timed CPU
baseline), never the product path.

Two layers:
  * the nine lib/pointnet2 operators, executed by oracle/pn2_oracle.c (a C restatement of the
    reference CUDA kernels with their exact fmaf order; built by oracle/Makefile);
  * the module wiring restated with plain PyTorch CPU ops, each function citing the reference
    lines it follows: QueryAndGroup, SharedMLP (eval), PointnetSAModuleVotes,
    PointnetFPModule, the Pointnet2Backbone composition driven directly by a
    state_dict with the reference's key names, and the situation re-encoding.

Parity pinning: the operators are pinned against the reference's own CUDA extension
(oracle/_ref, built unmodified from /root/reference by oracle/build_ref.py and run on the GPU
box; vectors in tests/golden/ref_cuda_ops.npz), the wiring against the reference's own Python
modules imported from /root/reference (tests/golden/ref_modules.npz, made by
tests/golden/make_ref_module_goldens.py).  The 1x1 conv / BatchNorm arithmetic itself is
third-party (PyTorch 1.12 + cuDNN 8.3 in the reference environment, environment.yml:60) and
is compared by tolerance only.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpn2_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
                os.path.join(_HERE, "pn2_oracle.c")):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.pn2o_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().pn2o_num_threads())


def _f32(t):
    t = torch.as_tensor(t)
    assert t.dtype == torch.float32 and not t.is_cuda
    return t.contiguous()


def _i32(t):
    t = torch.as_tensor(t)
    assert t.dtype == torch.int32 and not t.is_cuda
    return t.contiguous()


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


# ---- the nine operators (same names / argument order as pointnet2._ext) ------------------

def furthest_point_sampling(points, nsamples):
    points = _f32(points)
    B, N, _ = points.shape
    out = torch.zeros((B, nsamples), dtype=torch.int32)
    lib().pn2o_furthest_point_sampling(B, N, int(nsamples), _p(points), _p(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    B, m, _ = new_xyz.shape
    out = torch.zeros((B, m, nsample), dtype=torch.int32)
    lib().pn2o_ball_query(B, xyz.shape[1], m, ctypes.c_float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(out))
    return out


def three_nn(unknowns, knows):
    unknowns, knows = _f32(unknowns), _f32(knows)
    B, n, _ = unknowns.shape
    dist2 = torch.zeros((B, n, 3), dtype=torch.float32)
    idx = torch.zeros((B, n, 3), dtype=torch.int32)
    lib().pn2o_three_nn(B, n, knows.shape[1], _p(unknowns), _p(knows), _p(dist2), _p(idx))
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    B, c, m = points.shape
    n = idx.shape[1]
    out = torch.zeros((B, c, n), dtype=torch.float32)
    lib().pn2o_three_interpolate(B, c, m, n, _p(points), _p(idx), _p(weight), _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, c, n = grad_out.shape
    out = torch.zeros((B, c, m), dtype=torch.float32)
    lib().pn2o_three_interpolate_grad(B, c, n, int(m), _p(grad_out), _p(idx), _p(weight), _p(out))
    return out


def gather_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    m = idx.shape[1]
    out = torch.zeros((B, C, m), dtype=torch.float32)
    lib().pn2o_gather_points(B, C, N, m, _p(points), _p(idx), _p(out))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, m = grad_out.shape
    out = torch.zeros((B, C, n), dtype=torch.float32)
    lib().pn2o_gather_points_grad(B, C, int(n), m, _p(grad_out), _p(idx), _p(out))
    return out


def group_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    _, npoints, nsample = idx.shape
    out = torch.zeros((B, C, npoints, nsample), dtype=torch.float32)
    lib().pn2o_group_points(B, C, N, npoints, nsample, _p(points), _p(idx), _p(out))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, npoints, nsample = grad_out.shape
    out = torch.zeros((B, C, n), dtype=torch.float32)
    lib().pn2o_group_points_grad(B, C, int(n), npoints, nsample, _p(grad_out), _p(idx), _p(out))
    return out


# ---- module wiring restated with PyTorch CPU ops ---------------------------------------

class _CpuOps:
    """The C restatement above.  ``ops=`` lets bench.py run the SAME wiring over the reference's own
    CUDA extension (oracle/_ref) with CUDA tensors: that is the reference's stock GPU path."""
    furthest_point_sampling = staticmethod(furthest_point_sampling)
    ball_query = staticmethod(ball_query)
    three_nn = staticmethod(three_nn)
    three_interpolate = staticmethod(three_interpolate)
    gather_points = staticmethod(gather_points)
    group_points = staticmethod(group_points)


def query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz=True, normalize_xyz=False, ops=_CpuOps):
    """pointnet2_utils.py:317-376 (sample_uniformly off).  Returns (new_features, grouped_xyz, idx)."""
    ball_query, group_points = ops.ball_query, ops.group_points
    idx = ball_query(new_xyz, xyz, radius, nsample)                       # :334
    grouped_xyz = group_points(xyz.transpose(1, 2).contiguous(), idx)     # :347-348
    grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)     # :349
    if normalize_xyz:
        grouped_xyz = grouped_xyz / radius                                # :350-351
    if features is not None:
        grouped = group_points(features, idx)                             # :354
        new_features = torch.cat([grouped_xyz, grouped], dim=1) if use_xyz else grouped   # :355-360
    else:
        new_features = grouped_xyz
    return new_features, grouped_xyz, idx


def shared_mlp_eval(x, sd, prefix, nlayers, eps=1e-5):
    """pytorch_utils.py:11-36,67-121 in eval mode, from state-dict tensors:
    conv 1x1 (no bias when BN follows) -> BatchNorm2d(running stats) -> ReLU, per layer."""
    for i in range(nlayers):
        w = sd["%slayer%d.conv.weight" % (prefix, i)]
        b = sd.get("%slayer%d.conv.bias" % (prefix, i))
        x = F.conv2d(x, w, b)
        key = "%slayer%d.bn.bn." % (prefix, i)
        if key + "weight" in sd:
            x = F.batch_norm(x, sd[key + "running_mean"], sd[key + "running_var"], sd[key + "weight"],
                             sd[key + "bias"], training=False, eps=eps)
        x = F.relu(x)
    return x


def _count_layers(sd, prefix):
    n = 0
    while "%slayer%d.conv.weight" % (prefix, n) in sd:
        n += 1
    return n


def sa_module_votes(xyz, features, sd, prefix, npoint, radius, nsample, use_xyz=True, normalize_xyz=False,
                    inds=None, ops=_CpuOps):
    """PointnetSAModuleVotes.forward with max pooling (pointnet2_modules.py:210-277).
    Returns (new_xyz, new_features, inds, ball_idx)."""
    if inds is None:
        inds = ops.furthest_point_sampling(xyz, npoint)                    # :235
    new_xyz = ops.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()   # :238-240
    grouped, _, idx = query_and_group(xyz, new_xyz, features, radius, nsample, use_xyz, normalize_xyz, ops)
    x = shared_mlp_eval(grouped, sd, prefix + "mlp_module.", _count_layers(sd, prefix + "mlp_module."))   # :251
    x = F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)            # :259-262,272
    return new_xyz, x, inds, idx


def fp_module(unknown, known, unknow_feats, known_feats, sd, prefix, ops=_CpuOps):
    """PointnetFPModule.forward (pointnet2_modules.py:376-421)."""
    three_interpolate = ops.three_interpolate
    dist2, idx = ops.three_nn(unknown, known)
    dist = torch.sqrt(dist2)                                               # pointnet2_utils.py:142
    dist_recip = 1.0 / (dist + 1e-8)                                       # :400
    norm = torch.sum(dist_recip, dim=2, keepdim=True)
    weight = dist_recip / norm                                             # :401-402
    interpolated = three_interpolate(known_feats, idx, weight)             # :404-406
    x = torch.cat([interpolated, unknow_feats], dim=1) if unknow_feats is not None else interpolated   # :412-416
    x = shared_mlp_eval(x.unsqueeze(-1), sd, prefix + "mlp.", _count_layers(sd, prefix + "mlp."))      # :418-419
    return x.squeeze(-1)


BACKBONE_LAYERS = (   # (name, npoint, radius, nsample)
    ("sa1", 2048, 0.2, 64), ("sa2", 1024, 0.4, 32), ("sa3", 512, 0.8, 16), ("sa4", 256, 1.2, 16))


def backbone(point_clouds, sd, layers=BACKBONE_LAYERS, ops=_CpuOps):
    """Pointnet2Backbone.forward (composition of SURVEY.md 3.3 / 8a-0) from a state_dict with
    the reference's key names.  Returns the same dictionary keys as the product."""
    pc = _f32(point_clouds) if ops is _CpuOps else point_clouds
    xyz = pc[..., 0:3].contiguous()
    features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
    out = {}
    for name, npoint, radius, nsample in layers:
        xyz, features, inds, idx = sa_module_votes(xyz, features, sd, name + ".", npoint, radius, nsample,
                                                  use_xyz=True, normalize_xyz=True, ops=ops)
        out[name + "_xyz"], out[name + "_features"], out[name + "_inds"], out[name + "_ball_idx"] = \
            xyz, features, inds, idx
    f = fp_module(out["sa3_xyz"], out["sa4_xyz"], out["sa3_features"], out["sa4_features"], sd, "fp1.", ops)
    f = fp_module(out["sa2_xyz"], out["sa3_xyz"], out["sa2_features"], f, sd, "fp2.", ops)
    out["fp2_features"] = f
    out["fp2_xyz"] = out["sa2_xyz"]
    out["fp2_inds"] = out["sa1_inds"][:, 0:out["fp2_xyz"].shape[1]]
    return out


# ---- situation re-encoding --------------------------------------------------------------

def quaternions_to_rotation_matrices(q):
    """sqa_module.py:12-30 (xyzw, not normalised)."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.zeros((q.shape[0], 3, 3), dtype=q.dtype)
    R[:, 0, 0] = 1 - 2 * (y ** 2 + z ** 2); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x ** 2 + z ** 2); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x ** 2 + y ** 2)
    return R


def batch_rotation_vector_to_matrix(v):
    """sqa_module.py:33-64: Rodrigues, identity rows where |v| < 1e-6."""
    theta = torch.norm(v, dim=1, keepdim=True)
    eye = torch.eye(3).repeat(v.shape[0], 1, 1)
    small = theta < 1e-6
    if small.all():
        return eye
    u = v / theta
    zeros = torch.zeros(v.shape[0], 1)
    K = torch.cat([torch.cat([zeros, -u[:, 2:3], u[:, 1:2]], dim=1),
                   torch.cat([u[:, 2:3], zeros, -u[:, 0:1]], dim=1),
                   torch.cat([-u[:, 1:2], u[:, 0:1], zeros], dim=1)], dim=1).view(-1, 3, 3)
    th = theta.view(-1, 1, 1)
    R = eye + torch.sin(th) * K + (1 - torch.cos(th)) * torch.matmul(K, K)
    R[small.squeeze(1)] = eye[small.squeeze(1)]
    return R


def batch_matrix_function(s):
    """situation3d/utils/temp.py:42-80: (B,7) -> (B,4,4)."""
    x, y, z, w = s[:, 3], s[:, 4], s[:, 5], s[:, 6]
    M = torch.zeros((s.shape[0], 4, 4), dtype=s.dtype)
    M[:, 0, 0] = x * x - y * y - z * z + w * w; M[:, 0, 1] = 2 * (x * y - z * w); M[:, 0, 2] = 2 * (x * z + y * w)
    M[:, 1, 0] = 2 * (x * y + z * w); M[:, 1, 1] = -x * x + y * y - z * z + w * w; M[:, 1, 2] = 2 * (y * z - x * w)
    M[:, 2, 0] = 2 * (x * z - y * w); M[:, 2, 1] = 2 * (y * z + x * w); M[:, 2, 2] = -x * x - y * y + z * z + w * w
    M[:, 0, 3], M[:, 1, 3], M[:, 2, 3], M[:, 3, 3] = s[:, 0], s[:, 1], s[:, 2], 1
    return M


def reencode(tokens, positions, situation, w1, b1, w2, b2, to_agent_frame=False, sigma=0.16):
    """temp.py:86-97 transform, sqa_module.py:274-278,319-321 embedding + add, :328-336 prior."""
    M = batch_matrix_function(situation)
    if to_agent_frame:
        R, t = M[:, :3, :3], M[:, :3, 3]
        new_pos = torch.bmm(positions - t[:, None, :], R)          # rows: R^T (p - t)
    else:
        aug = torch.cat([positions, torch.ones_like(positions[..., :1])], dim=2)
        new_pos = torch.bmm(aug, M.transpose(1, 2))[:, :, :3]
    pe = F.linear(F.gelu(F.linear(new_pos[..., :2], w1, b1)), w2, b2)
    out = tokens + pe
    dist = torch.norm(positions[..., :2] - situation[:, None, :2], dim=2)
    wgt = torch.exp(-dist ** 2 / (2 * sigma ** 2))
    prior = wgt / wgt.sum(dim=1, keepdim=True)
    return out, new_pos, prior


# ---- visual-token construction ---------------------------------------
def column_tokens(coords_list, feats_list, tensor_stride, num_points=256, voxel_size=0.02):
    """situation3d/models/sqa_module.py:297-317 restated with the same PyTorch (CPU) calls in the same order, RNG
    calls included: unique (x, y) columns (:298-299), scatter_reduce_('mean') onto zeros, i.e. sum / (count + 1)
    (:300-301), randperm / randint sampling (:303-308), positions in metres (:311).
    Returns (scene_feat (B,T,C), scene_positions (B,T,2), [unique_coords], [reduced_feats], [inverse])."""
    scene_feat, scene_positions, uniq, red, inv = [], [], [], [], []
    for coords, feats in zip(coords_list, feats_list):
        reduced_coords = coords[:, [0, 1]]
        unique_coords, indices = reduced_coords.unique(dim=0, return_inverse=True)
        reduced_feats = torch.zeros(unique_coords.size(0), feats.size(1))
        reduced_feats = reduced_feats.scatter_reduce_(0, indices.unsqueeze(-1).expand_as(feats), feats, reduce='mean')
        if num_points < unique_coords.size(0):
            sampled = torch.randperm(unique_coords.size(0))[:num_points]
        else:
            sampled = torch.cat([torch.randperm(unique_coords.size(0)),
                                 torch.randint(0, unique_coords.size(0), (num_points - unique_coords.size(0),))])
        positions = (unique_coords[sampled] + torch.tensor(tensor_stride[0:2]) / 2) * voxel_size
        scene_feat.append(reduced_feats[sampled].unsqueeze(0))
        scene_positions.append(positions.unsqueeze(0))
        uniq.append(unique_coords); red.append(reduced_feats); inv.append(indices)
    return torch.cat(scene_feat, dim=0), torch.cat(scene_positions, dim=0), uniq, red, inv


# ---- voxel-coordinate positional embedding before the Q-Former ---------
def voxel_pe_table(rows=256, channels=1408 // 3, layout="concat"):
    """PositionalEncoding1D(channels) evaluated on zeros(1, rows, channels) and squeezed
    (3DLLM_BLIP2-base/lavis/models/blip2_models/blip2_t5.py:93-95).  The class comes from the third-party package
    ``positional_encodings`` (imported at blip2_t5.py:11; not vendored, no version pinned anywhere in the
    reference) -- PARITY UNPINNED for this function: it restates the package's published algorithm in float64 for
    both published row layouts (``cat(sin, cos)`` up to 5.x, interleaved sin/cos from 6.0) and is compared by
    tolerance only.  The gather/add below does not depend on it (the table is an input)."""
    ch = (channels + 1) // 2 * 2
    inv_freq = 1.0 / (10000.0 ** (np.arange(0, ch, 2, dtype=np.float64) / ch))
    ang = np.arange(rows, dtype=np.float64)[:, None] * inv_freq[None, :]
    if layout == "concat":
        emb = np.concatenate([np.sin(ang), np.cos(ang)], axis=-1)
    else:
        emb = np.stack([np.sin(ang), np.cos(ang)], axis=-1).reshape(rows, ch)
    return emb[:, :channels]


def voxel_pe_all_pcs(pc, table, channels):
    """blip2_t5.py:107-116 / blip2_opt.py:94-102 in numpy: per sample, the table rows of the x, y and z coordinate
    side by side in the first 3*S channels of a zero (P, channels) array.  ``pc`` is truncated toward zero first
    (``.long()``, :107); negative indices wrap and out-of-range ones raise IndexError, as torch's indexing does."""
    pc = np.trunc(np.asarray(pc, dtype=np.float64)).astype(np.int64) if np.asarray(pc).dtype.kind == "f" else np.asarray(pc, dtype=np.int64)
    table = np.asarray(table, dtype=np.float32)
    R, S = table.shape
    if ((pc[..., :3] < -R) | (pc[..., :3] >= R)).any():
        raise IndexError("index out of range for a table of %d rows" % R)
    B, P = pc.shape[:2]
    all_pcs = np.zeros((B, P, channels), dtype=np.float32)
    for j in range(B):
        all_pcs[j, :, :3 * S] = np.concatenate([table[pc[j, :, i]] for i in range(3)], axis=-1)
    return all_pcs


def voxel_pe(pc_feat, pc, table, mode="add", scale=0.01):
    """mode "add": pc_feat + 0.01 * all_pcs with the product rounded to fp32 first (blip2_t5.py:118);
    mode "cat": cat([pc_feat, all_pcs], 1) (blip2_opt.py:104)."""
    pc_feat = np.asarray(pc_feat, dtype=np.float32)
    all_pcs = voxel_pe_all_pcs(pc, table, pc_feat.shape[-1])
    if mode == "add":
        return pc_feat + np.float32(scale) * all_pcs
    return np.concatenate([pc_feat, all_pcs], axis=1)


# ---- point <-> pixel correspondences and back-projection ---------------
def corner_points(intrinsic, depth_min, depth_max, image_dims):
    """ProjectionHelper._compute_corner_points with depth_to_skeleton (lib/projection.py:17-21,29-46) -> (8,3) f32:
    the image corners (0,0), (W-1,0), (W-1,H-1), (0,H-1) at depth_min, then at depth_max, in the camera frame.
    Evaluated in fp32 tensor arithmetic as the reference does when `intrinsic` is a float tensor."""
    intrinsic = torch.as_tensor(intrinsic, dtype=torch.float32)
    out = torch.empty(8, 3)
    k = 0
    for depth in (depth_min, depth_max):
        for ux, uy in ((0, 0), (image_dims[0] - 1, 0), (image_dims[0] - 1, image_dims[1] - 1), (0, image_dims[1] - 1)):
            x = (ux - intrinsic[0][2]) / intrinsic[0][0]
            y = (uy - intrinsic[1][2]) / intrinsic[1][1]
            out[k] = torch.Tensor([depth * x, depth * y, depth])
            k += 1
    return out


def _projection_args(intrinsic, depth_min, depth_max, image_dims, accuracy):
    intrinsic = torch.as_tensor(intrinsic, dtype=torch.float32)
    intr4 = torch.tensor([intrinsic[0][0], intrinsic[1][1], intrinsic[0][2], intrinsic[1][2]], dtype=torch.float32)
    range3 = torch.tensor([depth_min, depth_max, accuracy], dtype=torch.float32)
    return intr4, range3, _f32(corner_points(intrinsic, depth_min, depth_max, image_dims))


def frustum_planes(camera_to_world, intrinsic, depth_min, depth_max, image_dims):
    """compute_frustum_corners + compute_frustum_normals (lib/projection.py:48-119) -> ((8,4), (6,3))."""
    c2w = _f32(camera_to_world)
    cp = _f32(corner_points(intrinsic, depth_min, depth_max, image_dims))
    corners, normals = torch.empty(8, 4), torch.empty(6, 3)
    lib().pn2o_frustum_planes(_p(c2w), _p(cp), _p(corners), _p(normals))
    return corners, normals


def compute_projection(points, depth, camera_to_world, world_to_camera, intrinsic, depth_min, depth_max, image_dims, accuracy):
    """ProjectionHelper.compute_projection (lib/projection.py:191-254) for one view: (indices_3d, indices_2d), both
    (num_points + 1,) int64 with the count first, or None where the reference returns None.  ``world_to_camera`` is
    the caller's ``torch.inverse(camera_to_world)`` (:203)."""
    points, depth = _f32(points), _f32(depth)
    c2w, w2c = _f32(camera_to_world), _f32(world_to_camera)
    intr4, range3, cp = _projection_args(intrinsic, depth_min, depth_max, image_dims, accuracy)
    n = points.shape[0]
    assert depth.numel() == image_dims[0] * image_dims[1]
    i3 = torch.empty(n + 1, dtype=torch.int64)
    i2 = torch.empty(n + 1, dtype=torch.int64)
    fn = lib().pn2o_compute_projection
    fn.restype = ctypes.c_longlong
    cnt = fn(ctypes.c_int(n), _p(points), _p(depth), _p(c2w), _p(w2c), _p(intr4), _p(range3), ctypes.c_int(image_dims[0]),
             ctypes.c_int(image_dims[1]), _p(cp), _p(i3), _p(i2))
    return None if cnt == 0 else (i3, i2)


def project(label, lin_indices_3d, lin_indices_2d, num_points):
    """ProjectionHelper.project (lib/projection.py:257-279) in numpy: (C, num_points) zeros with the listed points
    filled from the listed pixels."""
    label = np.asarray(label, dtype=np.float32)
    c = 1 if label.ndim == 2 else label.shape[0]
    i3, i2 = np.asarray(lin_indices_3d), np.asarray(lin_indices_2d)
    out = np.zeros((c, num_points), dtype=np.float32)
    k = int(i3[0])
    if k > 0:
        out[:, i3[1:1 + k]] = label.reshape(c, -1)[:, i2[1:1 + k]]
    return out


def points_in_frustum(corner_coords, normals, new_pts, return_mask=False):
    """ProjectionHelper.points_in_frustum (lib/projection.py:121-155): bool mask or the number of points inside."""
    cc = _f32(torch.as_tensor(corner_coords, dtype=torch.float32).reshape(8, 4))
    nr = _f32(torch.as_tensor(normals, dtype=torch.float32).reshape(6, 3))
    pts = _f32(new_pts)
    mask = torch.empty(pts.shape[0], dtype=torch.uint8)
    fn = lib().pn2o_points_in_frustum
    fn.restype = ctypes.c_longlong
    cnt = fn(ctypes.c_int(pts.shape[0]), _p(pts), _p(cc), _p(nr), _p(mask))
    return mask.bool() if return_mask else int(cnt)
