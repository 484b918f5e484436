"""The consumer of the visual tokens: ``SIG3D.scene_feat_linear``.

``SceneFeatLinear`` is ``nn.Sequential(nn.Linear(256, 768), nn.GELU())`` exactly as the reference declares it
(situation3d/models/sqa_module.py:180-183: parameter names ``0.weight`` / ``0.bias``, so a SIG3D checkpoint's
``scene_feat_linear.*`` loads unchanged); in eval mode on CUDA the linear layer, the bias and the exact-erf GELU run as
one tcgen05 kernel (csrc/head_tc.cu: bf16 operands, fp32 accumulation and output).  Training, or widths the kernel
does not cover, run the two PyTorch modules as the reference does.  Also here: the VoteNet-style ``seed_*`` aliases
of the backbone outputs that ``lib/loss_helper.py:47-59`` reads.
"""
import torch
import torch.nn as nn

from ._lib import check, lib, ptr, stream_ptr


class SceneFeatLinear(nn.Sequential):
    def __init__(self, in_dim=256, hidden_size=768, precision="bf16"):
        super().__init__(nn.Linear(in_dim, hidden_size), nn.GELU())
        self.precision = precision
        self._key, self._image = None, None

    def _weight_image(self):
        w = self[0].weight
        key = (w.data_ptr(), w._version, w.device)
        if key != self._key:
            n, k = w.shape
            image = torch.empty(lib.pn2_linear_gelu_tc_weight_image_bytes(k, n), dtype=torch.uint8, device=w.device)
            with torch.cuda.device(w.device):
                check(lib.pn2_linear_gelu_tc_pack_weights(k, n, ptr(w.detach().float().contiguous()), ptr(image), stream_ptr()),
                      "linear_gelu_tc_pack_weights")
            self._key, self._image = key, image
        return self._image

    def forward(self, x):
        lin = self[0]
        n, k = lin.weight.shape
        fused = (not self.training and self.precision == "bf16" and x.is_cuda and x.dtype == torch.float32
                 and lin.bias is not None and not (torch.is_grad_enabled() and (x.requires_grad or lin.weight.requires_grad))
                 and lib.pn2_linear_gelu_tc_supported(k, n))
        if not fused:
            if not x.is_cuda:
                raise RuntimeError("SceneFeatLinear: CUDA tensor required (this package has no CPU path)")
            return super().forward(x)       # training / autograd, fp32 arm or widths outside the kernel: the reference's modules
        x2 = x.contiguous().view(-1, k)
        out = torch.empty((x2.shape[0], n), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.pn2_linear_gelu_tc_forward(x2.shape[0], k, n, ptr(x2), ptr(self._weight_image()),
                                                 ptr(lin.bias.detach().float().contiguous()), ptr(out), stream_ptr()),
                  "linear_gelu_tc_forward")
        return out.view(*x.shape[:-1], n)


def add_seed_aliases(data_dict):
    """VoteNet naming of the backbone outputs (``seed_xyz`` / ``seed_features`` / ``seed_inds`` = ``fp2_*``), the keys
    ``lib/loss_helper.py:47-59`` and the vote / proposal stages of the upstream pipeline read."""
    data_dict["seed_xyz"] = data_dict["fp2_xyz"]
    data_dict["seed_features"] = data_dict["fp2_features"]
    data_dict["seed_inds"] = data_dict["fp2_inds"]
    return data_dict
