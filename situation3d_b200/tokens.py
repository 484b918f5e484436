"""Visual-token construction in front of the situation re-encoding.

Mirrors the per-scene loop of ``SIG3D.forward`` (situation3d/models/sqa_module.py:297-315): the sparse encoder's
bottleneck voxels are pooled over z into unique (x, y) columns (ascending order, feature = sum / (count + 1):
``scatter_reduce_('mean')`` onto zeros counts the zero), ``num_points`` columns are drawn with the reference's own
RNG calls (``torch.randperm`` / ``torch.randint`` on the CPU generator, in the reference's order, so the same seed
draws the same tokens), and the tokens come back with their positions in metres.  One pooling launch for the whole
batch and one gather launch replace the loop's ``unique`` / ``scatter_reduce_`` / indexing kernels per scene.
"""
import torch

from ._lib import check, lib, ptr, stream_ptr


def column_pool(coords_list, feats_list):
    """Per scene: unique (x, y) columns and their pooled features.
    Returns (offsets (B+1,) int32 cuda, ncols list[int], pooled_coords (M,2) int32, pooled_feats (M,C) f32,
    inverse (M,) int32); scene s owns rows offsets[s] .. offsets[s] + ncols[s] of the pooled arrays."""
    B = len(coords_list)
    dev = feats_list[0].device
    sizes = [int(c.shape[0]) for c in coords_list]
    if max(sizes) > lib.pn2_column_pool_max_voxels():
        raise RuntimeError("column_pool: more than %d voxels in a scene" % lib.pn2_column_pool_max_voxels())
    coords = torch.cat([c[:, :3].to(torch.int32) for c in coords_list]).contiguous()
    feats = torch.cat([f.to(torch.float32) for f in feats_list]).contiguous()
    if not (coords.is_cuda and feats.is_cuda):
        raise RuntimeError("column_pool: CUDA tensors required (there is no CPU path)")
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    offsets = torch.tensor(offs, dtype=torch.int32, device=dev)
    M, C = feats.shape
    ncols = torch.empty(B, dtype=torch.int32, device=dev)
    pooled_coords = torch.empty((M, 2), dtype=torch.int32, device=dev)
    pooled_feats = torch.empty((M, C), dtype=torch.float32, device=dev)
    inverse = torch.empty(M, dtype=torch.int32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.pn2_column_pool(B, ptr(offsets), ptr(coords), ptr(feats), C, ptr(ncols), ptr(pooled_coords),
                                  ptr(pooled_feats), ptr(inverse), ptr(status), stream_ptr()), "column_pool")
    host = torch.cat([ncols, status]).cpu()          # the reference synchronises here too (unique() returns a size)
    if int(host[-1]) != 0:
        raise RuntimeError("column_pool: voxel coordinates outside +-2^19")
    return offsets, [int(v) for v in host[:-1]], pooled_coords, pooled_feats, inverse


def column_tokens(coords_list, feats_list, tensor_stride, num_points=256, voxel_size=0.02):
    """(scene_feat (B,num_points,C), scene_positions (B,num_points,2)) as sqa_module.py:297-317 builds them."""
    offsets, ncols, pooled_coords, pooled_feats, _ = column_pool(coords_list, feats_list)
    dev = pooled_feats.device
    sampled = []
    for u in ncols:                                   # sqa_module.py:303-308, same RNG calls in the same order
        if num_points < u:
            idx = torch.randperm(u)[:num_points]
        else:
            idx = torch.cat([torch.randperm(u), torch.randint(0, u, (num_points - u,))])
        sampled.append(idx)
    sampled = torch.stack(sampled).to(torch.int32).to(dev)
    B, C = len(ncols), pooled_feats.shape[1]
    tokens = torch.empty((B, num_points, C), dtype=torch.float32, device=dev)
    positions = torch.empty((B, num_points, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.pn2_token_gather(B, num_points, C, ptr(offsets), ptr(sampled), ptr(pooled_feats), ptr(pooled_coords),
                                   float(tensor_stride[0]) / 2, float(tensor_stride[1]) / 2, float(voxel_size), ptr(tokens),
                                   ptr(positions), stream_ptr()), "token_gather")
    return tokens, positions
