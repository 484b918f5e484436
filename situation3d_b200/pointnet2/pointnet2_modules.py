"""PointNet++ set-abstraction and feature-propagation layers -- same class names,
keyword-only constructors, forward signatures, return values and state-dict keys as the
reference's lib/pointnet2/pointnet2_modules.py:

    _PointnetSAModuleBase :26-75     PointnetSAModuleMSG     :78-126    PointnetSAModule :129-161
    PointnetSAModuleVotes :164-277   PointnetSAModuleMSGVotes :279-358
    PointnetFPModule      :361-421   PointnetLFPModuleMSG    :423-501

Every module runs on this repository's CUDA operators.  In eval mode the two modules the
backbone is built from (PointnetSAModuleVotes with max pooling, PointnetFPModule) take the
fused path: sampling kernel -> ball query -> ONE kernel for neighbour gathering + shared MLP
+ max-pool (resp. 3-NN weights + interpolation + concat + MLP), with BatchNorm folded.  In
training mode, or for any configuration the fused kernels do not cover (avg/rbf pooling,
uniform resampling, non-ReLU stacks), they run operator by operator exactly as the reference
wires them, with autograd through the ``*_grad`` kernels.  ``fused=False`` forces that path.
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils
from . import pytorch_utils as pt_utils
from .. import fused as _fused


def _inference_only(module, *tensors):
    """True when nothing downstream can ask for gradients: the fused kernels are forward-only."""
    if module.training:
        return False
    if not torch.is_grad_enabled():
        return True
    if any(p.requires_grad for p in module.parameters()):
        return False
    return not any(t is not None and t.requires_grad for t in tensors)


def _all_f32(*tensors):
    """The fused kernels read raw float32 pointers; anything else (half / double after net.half(), autocast) takes the
    unfused path, whose ``_ext`` wrappers raise like the reference's CHECK_IS_FLOAT (utils.h:19-25)."""
    return all(t is None or t.dtype == torch.float32 for t in tensors)


def _max_over_samples(x):
    # (B, C, npoint, nsample) -> (B, C, npoint)
    return F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)


def _sample_centres(xyz, npoint, inds=None):
    """FPS (unless indices are given) and the gather of their coordinates: (inds, new_xyz)."""
    if npoint is None:
        return inds, None
    if inds is None:
        inds = pointnet2_utils.furthest_point_sample(xyz, npoint)
    flipped = xyz.transpose(1, 2).contiguous()
    new_xyz = pointnet2_utils.gather_operation(flipped, inds).transpose(1, 2).contiguous()
    return inds, new_xyz


def _make_groupers_and_mlps(npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly):
    groupers, stacks = nn.ModuleList(), nn.ModuleList()
    for radius, nsample, spec in zip(radii, nsamples, mlps):
        groupers.append(
            pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz, sample_uniformly=sample_uniformly)
            if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
        if use_xyz:
            spec[0] += 3   # in place, like the reference: the caller's list is modified
        stacks.append(pt_utils.SharedMLP(spec, bn=bn))
    return groupers, stacks


class _PointnetSAModuleBase(nn.Module):
    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None) -> (torch.Tensor, torch.Tensor):
        """xyz (B,N,3), features (B,C,N) -> new_xyz (B,npoint,3), new_features (B,sum_k mlps[k][-1],npoint)"""
        _, new_xyz = _sample_centres(xyz, self.npoint)
        pooled = [_max_over_samples(mlp(grouper(xyz, new_xyz, features)))
                  for grouper, mlp in zip(self.groupers, self.mlps)]
        return new_xyz, torch.cat(pooled, dim=1)


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Set abstraction with multi-scale grouping."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]],
                 bn: bool = True, use_xyz: bool = True, sample_uniformly: bool = False):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.groupers, self.mlps = _make_groupers_and_mlps(npoint, radii, nsamples, mlps, bn, use_xyz,
                                                            sample_uniformly)


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz)


class PointnetSAModuleVotes(nn.Module):
    """Set abstraction that also returns the sampled indices (VoteNet seeds).

    Extra keyword arguments over the reference: ``fused`` (default True) enables the fused
    eval-mode kernels, ``precision`` selects their arithmetic ("fp32" FFMA, or "bf16"
    tcgen05 tensor cores when built)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pooling: str = 'max', sigma: float = None,
                 normalize_xyz: bool = False, sample_uniformly: bool = False, ret_unique_cnt: bool = False,
                 fused: bool = True, precision: str = "fp32"):
        super().__init__()
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.pooling = pooling
        self.mlp_module = None
        self.use_xyz = use_xyz
        self.sigma = sigma
        if self.sigma is None:
            self.sigma = self.radius / 2
        self.normalize_xyz = normalize_xyz
        self.ret_unique_cnt = ret_unique_cnt
        self.sample_uniformly = sample_uniformly
        self.fused = fused
        self.precision = precision

        if npoint is not None:
            self.grouper = pointnet2_utils.QueryAndGroup(
                radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True, normalize_xyz=normalize_xyz,
                sample_uniformly=sample_uniformly, ret_unique_cnt=ret_unique_cnt)
        else:
            self.grouper = pointnet2_utils.GroupAll(use_xyz, ret_grouped_xyz=True)

        mlp_spec = mlp
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3   # in place (pointnet2_modules.py:204-206)
        self.mlp_module = pt_utils.SharedMLP(mlp_spec, bn=bn)
        self._image = _fused.MlpImage()

    # ---- fused eval path ---------------------------------------------------------------
    def _fused_image(self, xyz, features=None):
        ok = self.fused and xyz.is_cuda and _all_f32(xyz, features) and _inference_only(self, xyz, features)
        if not ok or self.npoint is None or self.pooling != 'max' or self.sample_uniformly:
            return None
        img = self._image.get(self.mlp_module, self.precision)
        return img if img is not None and img.image is not None else None

    def forward_rows(self, xyz, rows, ld, c, inds=None, img=None):
        """Fused layer on channel-last feature rows (``rows`` is a tensor whose first element is
        feature 0 of point (0,0); ``ld`` = row pitch in elements).  Returns
        (new_xyz, out (B,cout,npoint), out_rows (B,npoint,cout), inds)."""
        if img is None:
            img = self._fused_image(xyz)
        if inds is None:
            inds, new_xyz = _fused.fps_with_xyz(xyz, self.npoint)
        else:
            _, new_xyz = _sample_centres(xyz, self.npoint, inds)
        idx = _fused.ball_query(xyz, new_xyz, self.radius, self.nsample)
        inv_r = 1.0 / self.radius if self.normalize_xyz else 1.0
        out, out_rows = _fused.SA_FORWARD[self.precision](img, xyz, new_xyz, idx, rows, ld, c, self.use_xyz, inv_r)
        return new_xyz, out, out_rows, inds

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None,
                inds: torch.Tensor = None) -> (torch.Tensor, torch.Tensor):
        """xyz (B,N,3), features (B,C,N), inds (B,npoint) optional ->
        new_xyz (B,npoint,3), new_features (B,mlp[-1],npoint), inds (B,npoint) [, unique_cnt]"""
        if inds is not None:
            assert inds.shape[1] == self.npoint
        img = self._fused_image(xyz, features) if features is not None else None
        if img is not None:
            rows = _fused.rows_from_channels(features.contiguous())
            new_xyz, out, _, inds = self.forward_rows(xyz.contiguous(), rows, rows.shape[2], rows.shape[2], inds, img)
            return new_xyz, out, inds

        inds, new_xyz = _sample_centres(xyz, self.npoint, inds)
        grouped = self.grouper(xyz, new_xyz, features)
        if self.ret_unique_cnt:
            grouped_features, grouped_xyz, unique_cnt = grouped
        else:
            grouped_features, grouped_xyz = grouped
        new_features = self.mlp_module(grouped_features)   # (B, mlp[-1], npoint, nsample)

        if self.pooling == 'max':
            new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == 'avg':
            new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == 'rbf':
            # Gaussian-weighted sum over the samples, normalised by nsample (pointnet2_modules.py:267-271)
            rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (self.sigma ** 2) / 2)
            new_features = torch.sum(new_features * rbf.unsqueeze(1), -1, keepdim=True) / float(self.nsample)
        new_features = new_features.squeeze(-1)

        if self.ret_unique_cnt:
            return new_xyz, new_features, inds, unique_cnt
        return new_xyz, new_features, inds


class PointnetSAModuleMSGVotes(nn.Module):
    """Multi-scale set abstraction that also returns the sampled indices."""

    def __init__(self, *, mlps: List[List[int]], npoint: int, radii: List[float], nsamples: List[int],
                 bn: bool = True, use_xyz: bool = True, sample_uniformly: bool = False):
        super().__init__()
        assert len(mlps) == len(nsamples) == len(radii)
        self.npoint = npoint
        self.groupers, self.mlps = _make_groupers_and_mlps(npoint, radii, nsamples, mlps, bn, use_xyz,
                                                            sample_uniformly)

    def forward(self, xyz: torch.Tensor, features: torch.Tensor = None,
                inds: torch.Tensor = None) -> (torch.Tensor, torch.Tensor):
        inds, new_xyz = _sample_centres(xyz, self.npoint, inds)
        pooled = [_max_over_samples(mlp(grouper(xyz, new_xyz, features)))
                  for grouper, mlp in zip(self.groupers, self.mlps)]
        return new_xyz, torch.cat(pooled, dim=1), inds


class PointnetFPModule(nn.Module):
    """Propagates the features of a coarse (known) set to a fine (unknown) set by inverse-distance
    3-NN interpolation, concatenates the fine set's own features and applies a shared MLP."""

    def __init__(self, *, mlp: List[int], bn: bool = True, fused: bool = True, precision: str = "fp32"):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)
        self.fused = fused
        self.precision = precision
        self._image = _fused.MlpImage()

    def _fused_image(self, ref, *feats):
        if not (self.fused and ref.is_cuda and _all_f32(ref, *feats) and _inference_only(self, *feats)):
            return None
        img = self._image.get(self.mlp, self.precision, kind="fp")
        return img if img is not None and img.image is not None else None

    def forward_rows(self, unknown, known, skip_rows, known_rows, img=None):
        """Fused layer on channel-last rows: skip_rows (B,n,C1) or None, known_rows (B,m,C2).
        Returns (out (B,cout,n), out_rows (B,n,cout))."""
        if img is None:
            img = self._fused_image(unknown)
        return _fused.fp_layer(self.precision, img, unknown, known, known_rows, skip_rows)

    def forward(self, unknown: torch.Tensor, known: torch.Tensor, unknow_feats: torch.Tensor,
                known_feats: torch.Tensor) -> torch.Tensor:
        """unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m) -> (B,mlp[-1],n)"""
        if known is not None:
            img = self._fused_image(unknown, unknow_feats, known_feats)
            if img is not None:
                skip_rows = None if unknow_feats is None else _fused.rows_from_channels(unknow_feats.contiguous())
                known_rows = _fused.rows_from_channels(known_feats.contiguous())
                out, _ = self.forward_rows(unknown.contiguous(), known.contiguous(), skip_rows, known_rows, img)
                return out
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            norm = torch.sum(dist_recip, dim=2, keepdim=True)
            weight = dist_recip / norm
            interpolated_feats = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated_feats = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))

        if unknow_feats is not None:
            new_features = torch.cat([interpolated_feats, unknow_feats], dim=1)   # (B, C2 + C1, n)
        else:
            new_features = interpolated_feats
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)


class PointnetLFPModuleMSG(nn.Module):
    """Learnable feature propagation: set-abstraction style grouping from (xyz1, features1) around
    the points xyz2, followed by a post-MLP on the concatenation with features2."""

    def __init__(self, *, mlps: List[List[int]], radii: List[float], nsamples: List[int], post_mlp: List[int],
                 bn: bool = True, use_xyz: bool = True, sample_uniformly: bool = False):
        super().__init__()
        assert len(mlps) == len(nsamples) == len(radii)
        self.post_mlp = pt_utils.SharedMLP(post_mlp, bn=bn)
        self.groupers, self.mlps = _make_groupers_and_mlps(0, radii, nsamples, mlps, bn, use_xyz, sample_uniformly)

    def forward(self, xyz2: torch.Tensor, xyz1: torch.Tensor, features2: torch.Tensor,
                features1: torch.Tensor) -> torch.Tensor:
        """xyz2 (B,N2,3), xyz1 (B,N1,3), features2 (B,C2,N2), features1 (B,C1,N1) -> (B,sum_k post,N2)"""
        outs = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            new_features = _max_over_samples(mlp(grouper(xyz1, xyz2, features1)))
            if features2 is not None:
                new_features = torch.cat([new_features, features2], dim=1)
            outs.append(self.post_mlp(new_features.unsqueeze(-1)))
        return torch.cat(outs, dim=1).squeeze(-1)
