"""Situation-conditioned re-encoding of visual tokens.

Mirrors, name for name, the pieces of the reference this path consists of:

    quaternions_to_rotation_matrices   situation3d/models/sqa_module.py:12-30   (defined, unused in the release)
    batch_rotation_vector_to_matrix    situation3d/models/sqa_module.py:33-64   (defined, unused in the release)
    batch_matrix_function              situation3d/utils/temp.py:42-80          (7-D situation -> 4x4)
    SIG3D.pos_embed                    situation3d/models/sqa_module.py:274-278 (Linear(2,128) -> GELU -> Linear(128,256))
    tokens + pos_embed(positions)      situation3d/models/sqa_module.py:319-321
    Gaussian location prior            situation3d/models/sqa_module.py:328-336

``SituationReencoder`` holds ``pos_embed`` under the same parameter names as ``SIG3D`` (so
``pos_embed.0.weight`` ... load from a SIG3D checkpoint) and runs transform + embedding + add
(+ prior) as one fused CUDA launch.  The default transform is the executable statement the
reference contains (temp.py: p' = R p + t); ``to_agent_frame=True`` applies the inverse
R^T (p - t), a documented spec decision, not reference behaviour.
"""
import torch
import torch.nn as nn

from ._lib import check, lib, ptr, stream_ptr


def _cuda_f32(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError("%s must be a float32 CUDA tensor" % name)
    return t.contiguous()


def _matrices(fn, src, width, shape, name):
    src = _cuda_f32(src, name)
    if src.dim() != 2 or src.size(1) != width:
        raise RuntimeError("%s must be (B, %d)" % (name, width))
    out = torch.empty((src.size(0),) + shape, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        check(fn(src.size(0), ptr(src), ptr(out), stream_ptr()), name)
    return out


def quaternions_to_rotation_matrices(quaternions):
    """(B,4) xyzw (scipy order) -> (B,3,3); the quaternion is not normalised."""
    return _matrices(lib.pn2_quaternions_to_rotation_matrices, quaternions, 4, (3, 3), "quaternions")


def batch_rotation_vector_to_matrix(batch_rot_vec):
    """(B,3) rotation vectors -> (B,3,3) by Rodrigues' formula; identity when |v| < 1e-6."""
    return _matrices(lib.pn2_rotation_vectors_to_matrices, batch_rot_vec, 3, (3, 3), "batch_rot_vec")


def batch_matrix_function(quats):
    """(B,7) = (tx,ty,tz,qx,qy,qz,qw) -> (B,4,4) homogeneous transforms [R|t]."""
    return _matrices(lib.pn2_situation_matrices, quats, 7, (4, 4), "quats")


def reencode_tokens(tokens, positions, situation, w1, b1, w2, b2, *, to_agent_frame=False, sigma=0.16,
                    want_prior=True):
    """tokens (B,T,D), positions (B,T,3) or (B,T,2), situation (B,7) ->
    (tokens + pos_embed(p'_xy) (B,T,D), p' (B,T,3), prior (B,T) or None)."""
    tokens = _cuda_f32(tokens, "tokens")
    positions = _cuda_f32(positions, "positions")
    situation = _cuda_f32(situation, "situation")
    if positions.size(-1) == 2:   # the release keeps only voxel-column xy (sqa_module.py:311)
        positions = torch.cat([positions, torch.zeros_like(positions[..., :1])], dim=-1).contiguous()
    B, T, D = tokens.shape
    H = w1.shape[0]
    if positions.shape != (B, T, 3) or situation.shape != (B, 7):
        raise RuntimeError("positions must be (B,T,3|2) and situation (B,7)")
    if w1.shape != (H, 2) or b1.shape != (H,) or w2.shape != (D, H) or b2.shape != (D,):
        raise RuntimeError("pos_embed weights must be Linear(2,H) and Linear(H,D)")
    out = torch.empty_like(tokens)
    new_pos = torch.empty_like(positions)
    prior = torch.empty((B, T), dtype=torch.float32, device=tokens.device) if want_prior else None
    w1, b1, w2, b2 = (_cuda_f32(t.detach(), "pos_embed weight") for t in (w1, b1, w2, b2))
    with torch.cuda.device(tokens.device):
        check(lib.pn2_reencode_forward(B, T, D, H, 1 if to_agent_frame else 0, float(sigma), ptr(tokens),
                                       ptr(positions), ptr(situation), ptr(w1), ptr(b1), ptr(w2), ptr(b2),
                                       ptr(out), ptr(new_pos), ptr(prior), stream_ptr()), "reencode_forward")
    return out, new_pos, prior


class SituationReencoder(nn.Module):
    """Visual-token re-encoding: situation transform of the token positions, positional
    embedding of the transformed xy, residual add, and the Gaussian location prior."""

    def __init__(self, hidden=128, dim=256, sigma=0.16, to_agent_frame=False):
        super().__init__()
        self.pos_embed = nn.Sequential(nn.Linear(2, hidden), nn.GELU(), nn.Linear(hidden, dim))
        self.sigma = sigma
        self.to_agent_frame = to_agent_frame

    def _forward_autograd(self, tokens, positions, situation):
        """Differentiable twin of pn2_reencode_forward (temp.py:42-97 transform, sqa_module.py:274-278,319-321 embedding
        + add, :328-336 prior) in plain PyTorch ops."""
        if positions.size(-1) == 2:
            positions = torch.cat([positions, torch.zeros_like(positions[..., :1])], dim=-1)
        t, q = situation[:, :3], situation[:, 3:]
        x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([x * x - y * y - z * z + w * w, 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), -x * x + y * y - z * z + w * w, 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), -x * x - y * y + z * z + w * w], dim=1).view(-1, 3, 3)
        if self.to_agent_frame:
            new_pos = torch.matmul(positions - t[:, None, :], R)                # R^T (p - t)
        else:
            new_pos = torch.matmul(positions, R.transpose(1, 2)) + t[:, None, :]   # R p + t
        out = tokens + self.pos_embed(new_pos[..., :2])
        d2 = (positions[..., :2] - t[:, None, :2]).pow(2).sum(-1)
        prior = torch.exp(-d2 / (2 * self.sigma ** 2))
        prior = prior / prior.sum(dim=1, keepdim=True)
        return out, new_pos, prior

    def forward(self, data_dict):
        """Reads ``scene_feat`` (B,T,D) [falls back to ``att_feat_pre``], ``scene_positions`` (B,T,3|2)
        and ``auxiliary_task`` (B,7); writes ``att_feat_pre`` (input tokens), ``scene_feat`` (re-encoded),
        ``scene_positions_agent`` and ``auxiliary_task_loc_gt`` (the prior), the names SIG3D.forward uses
        (sqa_module.py:316-336)."""
        tokens = data_dict["scene_feat"] if "scene_feat" in data_dict else data_dict["att_feat_pre"]
        needs_grad = torch.is_grad_enabled() and (tokens.requires_grad or any(p.requires_grad for p in self.parameters()))
        if self.training or needs_grad:
            # the fused kernel is forward-only: training (or any caller that differentiates) runs the same statements
            # as PyTorch ops so that pos_embed, the tokens and everything upstream receive gradients
            out, new_pos, prior = self._forward_autograd(tokens, data_dict["scene_positions"], data_dict["auxiliary_task"])
            data_dict["att_feat_pre"] = tokens
            data_dict["scene_feat"] = out
            data_dict["scene_positions_agent"] = new_pos
            data_dict["auxiliary_task_loc_gt"] = prior
            return data_dict
        out, new_pos, prior = reencode_tokens(
            tokens, data_dict["scene_positions"], data_dict["auxiliary_task"],
            self.pos_embed[0].weight, self.pos_embed[0].bias, self.pos_embed[2].weight, self.pos_embed[2].bias,
            to_agent_frame=self.to_agent_frame, sigma=self.sigma)
        data_dict["att_feat_pre"] = tokens
        data_dict["scene_feat"] = out
        data_dict["scene_positions_agent"] = new_pos
        data_dict["auxiliary_task_loc_gt"] = prior
        return data_dict
