"""Pointnet2Backbone: SA1-SA4 + FP1-FP2 over the drop-in pointnet2 modules.

The class is absent from /root/reference; the composition is the upstream
VoteNet / ScanQA ``models/backbone_module.py`` that SIG3D's loss helpers still expect
(``lib/loss_helper.py:47-59`` reads ``fp2_xyz``/``fp2_features``/``fp2_inds`` as seeds), with
the layer table of SURVEY.md 8a-0 (corroborated by the shape comment at
``lib/pointnet2/pointnet2_modules.py:253-255``).

Eval-mode forward on CUDA is the fused B200 path:
  * the sampling pyramid (the four furthest-point chains, each fed by the previous level's
    sampled coordinates) depends on xyz only and runs ahead on its own stream;
  * per level, on the main stream: ball query -> one fused gather + MLP + max-pool kernel that
    reads channel-last rows (for SA1 straight out of ``point_clouds``) and writes both the
    reference's (B,C,npoint) tensor and the channel-last rows the next level gathers from;
  * FP1/FP2: three_nn -> one fused interpolate + concat + MLP kernel.
Training mode (or fused=False) runs the same modules operator by operator with autograd.
"""
import torch
import torch.nn as nn

from . import fused as _fused
from .pointnet2.pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes
from .streams import sampling_stream


class Pointnet2Backbone(nn.Module):
    r"""
    Parameters
    ----------
    input_feature_dim : int
        Channels per point besides xyz (129 for SIG3D/ScanQA: height + 128-d multiview).
    precision : "fp32" | "bf16"
        Arithmetic of the fused shared MLPs.
    """

    def __init__(self, input_feature_dim=0, *, precision="fp32", fused=True,
                 npoints=(2048, 1024, 512, 256), radii=(0.2, 0.4, 0.8, 1.2), nsamples=(64, 32, 16, 16)):
        super().__init__()
        self.input_feature_dim = input_feature_dim
        self.precision = precision
        self.fused = fused
        kw = dict(use_xyz=True, normalize_xyz=True, fused=fused, precision=precision)
        self.sa1 = PointnetSAModuleVotes(npoint=npoints[0], radius=radii[0], nsample=nsamples[0],
                                         mlp=[input_feature_dim, 64, 64, 128], **kw)
        self.sa2 = PointnetSAModuleVotes(npoint=npoints[1], radius=radii[1], nsample=nsamples[1],
                                         mlp=[128, 128, 128, 256], **kw)
        self.sa3 = PointnetSAModuleVotes(npoint=npoints[2], radius=radii[2], nsample=nsamples[2],
                                         mlp=[256, 128, 128, 256], **kw)
        self.sa4 = PointnetSAModuleVotes(npoint=npoints[3], radius=radii[3], nsample=nsamples[3],
                                         mlp=[256, 128, 128, 256], **kw)
        self.fp1 = PointnetFPModule(mlp=[256 + 256, 256, 256], fused=fused, precision=precision)
        self.fp2 = PointnetFPModule(mlp=[256 + 256, 256, 256], fused=fused, precision=precision)
        # training-mode layout: "rows" = the same mathematics on channel-last rows (train_rows.py: row gathers, one GEMM
        # per 1x1 convolution, BatchNorm over the row axis); "reference" = operator by operator as the reference wires it
        self.train_layout = "rows"
        self._side_streams = {}
        self.sm_partition = None      # optional streams.SmPartition: sampling side streams come from its FPS group

    @staticmethod
    def _break_up_pc(pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def set_precision(self, precision):
        self.precision = precision
        for m in (self.sa1, self.sa2, self.sa3, self.sa4, self.fp1, self.fp2):
            m.precision = precision

    # ---- compact transport format ------------------------------------------------------------
    def feature_row_elems(self):
        """bf16 elements per feature row of the compact input format: the features sit at elements 3 .. 3+C-1, as in
        an fp32 ``point_clouds`` row (elements 0..2 are zero), padded to whole 16-byte pieces."""
        return (3 + self.input_feature_dim + 7) // 8 * 8

    def pack_point_clouds(self, point_clouds):
        """(B, N, 3+C) f32 -> (xyz (B,N,3) f32, feature rows (B,N,row_elems) bf16): the format a data loader ships when
        host-to-device bandwidth bounds the step (half the bytes).  Works on CPU (pinned staging) and CUDA tensors;
        round-to-nearest-even, the same conversion the bf16 arm applies on the device."""
        C, R = self.input_feature_dim, self.feature_row_elems()
        xyz = point_clouds[..., :3].contiguous()
        rows = torch.zeros(point_clouds.shape[:2] + (R,), dtype=torch.bfloat16, device=point_clouds.device)
        rows[..., 3:3 + C] = point_clouds[..., 3:3 + C].to(torch.bfloat16)
        return xyz, rows

    # ---- fused eval path -------------------------------------------------------------------
    def _fused_images(self, pc):
        if not (self.fused and pc.is_cuda and pc.dtype == torch.float32 and (pc.size(-1) > 3 or self.input_feature_dim > 0)):
            return None
        sas = (self.sa1, self.sa2, self.sa3, self.sa4)
        imgs = [m._fused_image(pc) for m in sas] + [m._fused_image(pc) for m in (self.fp1, self.fp2)]
        return None if any(i is None for i in imgs) else imgs

    def _forward_fused(self, pc, imgs, data_dict, xyz_in=None, rows_bf16=None):
        """``pc`` (B, N, 3+C) f32, or -- the compact transport format of ``pack_point_clouds`` -- ``xyz_in`` (B, N, 3)
        f32 plus ``rows_bf16`` (B, N, row_elems) bf16 feature rows."""
        sas = (self.sa1, self.sa2, self.sa3, self.sa4)
        if pc is not None:
            B, N, W = pc.shape
            pc = pc.contiguous()
            dev = pc.device
            # the contiguous (B, N, 3) coordinates are written by the first sampling kernel itself while it loads the
            # points out of point_clouds (no separate copy kernel); everything that reads them runs after that kernel
            xyz = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
        else:
            B, N, _ = xyz_in.shape
            xyz, dev = xyz_in.contiguous(), xyz_in.device
            rows_bf16 = rows_bf16.contiguous()
        main = torch.cuda.current_stream(dev)
        # one sampling stream per caller stream: callers that keep several batches in flight (one stream
        # per batch) get independent sampling pyramids that overlap with each other's MLP kernels
        key = (dev.index, main.cuda_stream)
        side = self._side_streams.get(key)
        if side is None:
            part = self.sm_partition
            side = part.stream(part.FPS) if part is not None else sampling_stream(dev)
            self._side_streams[key] = side

        # buffers are allocated on the main stream; the side stream only fills them
        inds = [torch.empty((B, m.npoint), dtype=torch.int32, device=dev) for m in sas]
        cxyz = [torch.empty((B, m.npoint, 3), dtype=torch.float32, device=dev) for m in sas]
        ready = [torch.cuda.Event() for _ in sas]
        side.wait_stream(main)
        with torch.cuda.stream(side):
            src = xyz
            for lvl, m in enumerate(sas):
                if lvl == 0 and pc is not None:
                    _fused.fps_rows_into(pc, inds[0], cxyz[0], xyz)
                else:
                    _fused.fps_into(src, inds[lvl], cxyz[lvl])
                ready[lvl].record(side)
                src = cxyz[lvl]

        feats, rows = [], []
        if pc is not None:
            src_xyz, table, ld, c, skip = xyz, pc[..., 3:], W, W - 3, 3
        else:
            src_xyz, table, ld, c, skip = xyz, rows_bf16, rows_bf16.shape[2], self.input_feature_dim, 3
        # SA1's ball-query grid needs the scene only, not the centres: built here, it runs under the first
        # sampling kernel instead of after it
        grid0 = _fused.ball_query_grid_build(pc if pc is not None else xyz, N, sas[0].npoint, sas[0].radius,
                                             sas[0].nsample)
        for lvl, m in enumerate(sas):
            main.wait_event(ready[lvl])
            if lvl == 0 and grid0 is not None:
                idx = _fused.ball_query_grid_query(grid0, N, cxyz[0], m.radius, m.nsample)
            else:
                idx = _fused.ball_query(src_xyz, cxyz[lvl], m.radius, m.nsample)
            inv_r = 1.0 / m.radius if m.normalize_xyz else 1.0
            out, out_rows = _fused.SA_FORWARD[self.precision](imgs[lvl], src_xyz, cxyz[lvl], idx, table, ld, c,
                                                             m.use_xyz, inv_r, raw_skip=skip)
            feats.append(out)
            rows.append(out_rows)
            src_xyz, table, ld, c, skip = cxyz[lvl], out_rows, out_rows.shape[2], out_rows.shape[2], 0
            data_dict["sa%d_inds" % (lvl + 1)] = inds[lvl]
            data_dict["sa%d_xyz" % (lvl + 1)] = cxyz[lvl]
            data_dict["sa%d_features" % (lvl + 1)] = out

        _, fp1_rows = _fused.fp_layer(self.precision, imgs[4], cxyz[2], cxyz[3], rows[3], rows[2])
        fp2, _ = _fused.fp_layer(self.precision, imgs[5], cxyz[1], cxyz[2], fp1_rows, rows[1], want_rows=False)
        data_dict["fp2_features"] = fp2
        data_dict["fp2_xyz"] = cxyz[1]
        data_dict["fp2_inds"] = inds[0][:, 0:cxyz[1].shape[1]]
        return data_dict

    def prefetch_sampling(self, point_clouds):
        """Training: start the FPS chain of a FUTURE batch now, on the sampling side stream (it needs the coordinates
        only).  Called after this step's forward and before its backward, the next batch's 1.9 ms of dependent sampling
        rounds run under the backward pass; the next ``forward`` on that same tensor picks the result up."""
        if self.training and self.train_layout == "rows" and point_clouds.is_cuda and point_clouds.dtype == torch.float32:
            from . import train_rows
            train_rows._PREFETCHED[self] = train_rows.SamplingPyramid(self, point_clouds.contiguous())

    def forward(self, data_dict):
        r"""Reads ``data_dict["point_clouds"]`` (B, N, 3 + input_feature_dim) and adds
        ``sa{1..4}_{xyz,features,inds}``, ``fp2_features`` (B,256,1024), ``fp2_xyz``, ``fp2_inds``."""
        if "point_clouds" not in data_dict and "feature_rows_bf16" in data_dict:
            # compact transport format (pack_point_clouds): fp32 coordinates + bf16 feature rows.  The bf16 arm rounds
            # the features to bf16 as its first act anyway, so the results are bit-identical to the fp32 input's.
            xyz_in, rows = data_dict["xyz"], data_dict["feature_rows_bf16"]
            if self.training or self.precision != "bf16" or not self.fused:
                raise RuntimeError("feature_rows_bf16 input is the eval-mode bf16 fused path only")
            if not (xyz_in.is_cuda and xyz_in.dtype == torch.float32 and rows.dtype == torch.bfloat16 and
                    rows.shape[2] == self.feature_row_elems()):
                raise RuntimeError("xyz must be CUDA float32 (B,N,3) and feature_rows_bf16 CUDA bfloat16 (B,N,%d)"
                                   % self.feature_row_elems())
            imgs = self._fused_images(xyz_in)
            if imgs is None or imgs[0].f32_only:
                raise RuntimeError("the tensor-core path does not cover this configuration")
            return self._forward_fused(None, imgs, data_dict, xyz_in, rows)
        pc = data_dict["point_clouds"]
        if not self.training and not (torch.is_grad_enabled() and pc.requires_grad):
            imgs = self._fused_images(pc)
            if imgs is not None:
                return self._forward_fused(pc, imgs, data_dict)

        if self.training and self.train_layout == "rows" and pc.is_cuda and pc.dtype == torch.float32:
            from . import train_rows
            mods = (self.sa1, self.sa2, self.sa3, self.sa4)
            if all(train_rows.rows_supported(m.mlp_module) and m.pooling == "max" and not m.sample_uniformly and
                   not m.ret_unique_cnt for m in mods) and all(train_rows.rows_supported(m.mlp) for m in (self.fp1, self.fp2)):
                return train_rows.backbone_forward_rows(self, pc, data_dict)

        xyz, features = self._break_up_pc(pc)
        for lvl, m in enumerate((self.sa1, self.sa2, self.sa3, self.sa4), start=1):
            xyz, features, fps_inds = m(xyz, features)
            data_dict["sa%d_inds" % lvl] = fps_inds
            data_dict["sa%d_xyz" % lvl] = xyz
            data_dict["sa%d_features" % lvl] = features

        features = self.fp1(data_dict["sa3_xyz"], data_dict["sa4_xyz"], data_dict["sa3_features"],
                            data_dict["sa4_features"])
        features = self.fp2(data_dict["sa2_xyz"], data_dict["sa3_xyz"], data_dict["sa2_features"], features)
        data_dict["fp2_features"] = features
        data_dict["fp2_xyz"] = data_dict["sa2_xyz"]
        num_seed = data_dict["fp2_xyz"].shape[1]
        data_dict["fp2_inds"] = data_dict["sa1_inds"][:, 0:num_seed]
        return data_dict
