"""Backbone + situation re-encoding as one module (BASELINE.json config 3).

``SituatedSceneEncoder`` turns ``point_clouds`` (B, N, 3+C) and a 7-D situation vector
``auxiliary_task`` (B, 7) = (tx, ty, tz, qx, qy, qz, qw) into the visual tokens SIG3D's fusion stage
consumes: the first ``num_tokens`` seeds of the backbone (an FPS prefix, so they cover the scene), their
256-d ``fp2_features`` as tokens, re-encoded by ``SituationReencoder`` (position transform + positional
embedding + location prior).  The dictionary keys are the ones ``SIG3D.forward`` uses
(``situation3d/models/sqa_module.py:316-336``: ``att_feat_pre``, ``scene_positions``,
``auxiliary_task_loc_gt``), so the result can replace the MinkUNet branch of the release
(``sqa_module.py:283-321``) in front of ``scene_feat_linear`` (``:344``).
"""
import torch.nn as nn

from .backbone_module import Pointnet2Backbone
from .heads import SceneFeatLinear, add_seed_aliases
from .reencode import SituationReencoder


class SituatedSceneEncoder(nn.Module):
    def __init__(self, input_feature_dim=129, num_tokens=256, *, precision="bf16", to_agent_frame=False, sigma=0.16,
                 hidden_size=None):
        super().__init__()
        self.backbone_net = Pointnet2Backbone(input_feature_dim=input_feature_dim, precision=precision)
        self.reencoder = SituationReencoder(hidden=128, dim=256, sigma=sigma, to_agent_frame=to_agent_frame)
        self.num_tokens = num_tokens
        # SIG3D.scene_feat_linear (sqa_module.py:180-183,344): present when hidden_size is given (768 in the release)
        self.scene_feat_linear = SceneFeatLinear(256, hidden_size, precision) if hidden_size else None

    def forward(self, data_dict):
        data_dict = self.backbone_net(data_dict)
        t = self.num_tokens
        data_dict["scene_feat"] = data_dict["fp2_features"][:, :, :t].transpose(1, 2).contiguous()   # (B, t, 256)
        data_dict["scene_positions"] = data_dict["fp2_xyz"][:, :t].contiguous()                       # (B, t, 3)
        data_dict["scene_token_inds"] = data_dict["fp2_inds"][:, :t]
        add_seed_aliases(data_dict)
        data_dict = self.reencoder(data_dict)
        if self.scene_feat_linear is not None:
            data_dict["scene_feat_hidden"] = self.scene_feat_linear(data_dict["scene_feat"])      # (B, t, hidden_size)
        return data_dict
