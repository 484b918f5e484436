"""Positional embedding looked up by integer voxel coordinate, in front of the Q-Former.

Mirrors the second "PE before Q-Former" site of the reference, the 3D-LLM BLIP-2 wrappers:

* ``Blip2T5.__init__`` (3DLLM_BLIP2-base/lavis/models/blip2_models/blip2_t5.py:93-95) builds a (256, 469) sinusoid
  table with ``positional_encodings.torch_encodings.PositionalEncoding1D(1408 // 3)``;
* ``Blip2T5.forward`` / ``predict_answers`` (:104-118, :279-293) index it with the x, y, z voxel coordinate of every
  point in a per-sample CPU loop, fill channels 0..1406 of a zero CPU tensor, copy that to the GPU and add
  ``0.01 *`` it to ``pc_feat``;
* ``Blip2OPT.forward`` (blip2_opt.py:92-104) builds the same tensor and concatenates it to ``pc_feat`` along the
  point axis instead.

Here the loop, the host tensor and its copy are one kernel launch (csrc/voxel_pe.cu) that streams ``pc_feat`` once.

``positional_encodings`` is a third-party package the reference neither vendors nor pins.  Its published
``PositionalEncoding1D`` changed layout between releases: up to 5.x the row is ``cat(sin, cos)``, from 6.0 on sin and
cos are interleaved.  ``sinusoid_table`` restates both (``layout=``); which one a checkpoint was trained with is a
property of the user's environment, so the table can also be handed in (``VoxelPositionalEmbedding(table=...)``).
"""
import torch
from torch import nn

from ._lib import check, lib, ptr, stream_ptr

_COORD_KIND = {torch.int32: 0, torch.int64: 1, torch.float32: 2}


def sinusoid_table(rows=256, channels=1408 // 3, layout="concat"):
    """``PositionalEncoding1D(channels)(zeros(1, rows, channels)).squeeze()`` (blip2_t5.py:93-95), on the CPU in
    fp32 with the package's own operation order: inv_freq = 1 / 10000^(arange(0, ch', 2) / ch') with ch' = channels
    rounded up to even, angle = outer(position, inv_freq), then sin / cos laid out per ``layout`` and cut to
    ``channels`` columns.  Runs once at construction, like the reference's."""
    ch = (channels + 1) // 2 * 2
    inv_freq = 1.0 / (10000 ** (torch.arange(0, ch, 2).float() / ch))
    pos = torch.arange(rows).type(inv_freq.type())
    ang = torch.einsum("i,j->ij", pos, inv_freq)
    if layout == "concat":            # positional_encodings <= 5.x
        emb = torch.cat((ang.sin(), ang.cos()), dim=-1)
    elif layout == "interleave":      # positional_encodings >= 6.0
        emb = torch.flatten(torch.stack((ang.sin(), ang.cos()), dim=-1), -2, -1)
    else:
        raise ValueError("layout must be 'concat' or 'interleave'")
    return emb[:, :channels].contiguous()


def voxel_pe(pc_feat, pc, table, mode="add", scale=0.01, out=None, validate=True):
    """``pc_feat`` (B,P,C) f32 cuda, ``pc`` (B,P,>=3) int32 / int64 / float32 voxel coordinates (floats are truncated
    like the reference's ``.long()``), ``table`` (R,S) f32 with 3*S <= C.
    mode "add": returns pc_feat + scale * all_pcs (blip2_t5.py:118); mode "cat": returns cat([pc_feat, all_pcs], 1)
    (blip2_opt.py:104).  A coordinate outside [-R, R) raises IndexError as the reference's indexing does; that
    reads one flag back (a sync the reference's CPU loop has anyway) -- ``validate=False`` skips it, e.g. under
    CUDA-graph capture, and returns (out, status) so that the flag can be inspected later."""
    if not (pc_feat.is_cuda and pc.is_cuda and table.is_cuda):
        raise RuntimeError("voxel_pe: CUDA tensors required (there is no CPU path)")
    if pc_feat.dtype != torch.float32 or table.dtype != torch.float32 or pc.dtype not in _COORD_KIND:
        raise RuntimeError("voxel_pe: pc_feat / table must be float32, pc int32, int64 or float32")
    if pc_feat.dim() != 3 or pc.dim() != 3 or pc.shape[:2] != pc_feat.shape[:2] or pc.shape[2] < 3 or table.dim() != 2:
        raise RuntimeError("voxel_pe: expected pc_feat (B,P,C), pc (B,P,>=3), table (R,S)")
    B, P, C = pc_feat.shape
    R, S = table.shape
    if 3 * S > C:
        raise RuntimeError("voxel_pe: 3 x %d table columns do not fit %d channels" % (S, C))   # the reference's slice assignment fails too
    if mode not in ("add", "cat"):
        raise ValueError("mode must be 'add' or 'cat'")
    pc_feat, pc, table = pc_feat.contiguous(), pc.contiguous(), table.contiguous()
    shape = (B, P, C) if mode == "add" else (B, 2 * P, C)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=pc_feat.device)
    elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != pc_feat.device:
        raise RuntimeError("voxel_pe: out must be a contiguous float32 tensor of shape %s" % (shape,))
    status = torch.empty(1, dtype=torch.int32, device=pc_feat.device)
    with torch.cuda.device(pc_feat.device):
        check(lib.pn2_voxel_pe(B, P, C, S, R, _COORD_KIND[pc.dtype], pc.shape[2], ptr(pc), ptr(table), ptr(pc_feat), ptr(out),
                               float(scale), 0 if mode == "add" else 1, ptr(status), stream_ptr()), "voxel_pe")
    if not validate:
        return out, status
    if int(status.item()) != 0:
        raise IndexError("voxel_pe: voxel coordinate out of range for a table of %d rows" % R)
    return out


class VoxelPositionalEmbedding(nn.Module):
    """Holds ``pos_embedding`` (the name the reference gives the table, blip2_t5.py:95) and applies it.

        pe = VoxelPositionalEmbedding().cuda()
        pc_embeds = pe(samples["pc_feat"], samples["pc"])                  # Blip2T5:  + 0.01 * all_pcs
        pc_embeds = pe(samples["pc_feat"], samples["pc"], mode="cat")      # Blip2OPT: cat along points
    """

    def __init__(self, rows=256, channels=1408 // 3, layout="concat", table=None, scale=0.01):
        super().__init__()
        if table is None:
            table = sinusoid_table(rows, channels, layout)
        self.register_buffer("pos_embedding", table.detach().to(torch.float32).contiguous(), persistent=False)
        self.scale = scale

    def forward(self, pc_feat, pc, mode="add", out=None):
        return voxel_pe(pc_feat, pc, self.pos_embedding, mode=mode, scale=self.scale, out=out)
