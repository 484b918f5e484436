"""Workload of bench.py's reference arm, built WITHOUT importing the product package.  TEST INFRASTRUCTURE.

  * scenes: the synthetic ScanNet-shaped generator (situation3d_b200/synthetic.py, numpy only) is loaded by file
    path, so ``situation3d_b200`` never enters sys.modules and libpn2_b200.so is never loaded by this arm;
  * weights: a Pointnet2Backbone state_dict with the reference's key names, initialised like the
    reference's modules (kaiming_normal_ conv weights, pytorch_utils.py:96; BatchNorm with randomised running
    statistics and affine parameters so that eval-mode folding is non-trivial).
"""
import importlib.util
import os

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# SharedMLP widths after the +3 xyz channels
SA_MLPS = {"sa1": (64, 64, 128), "sa2": (128, 128, 256), "sa3": (128, 128, 256), "sa4": (128, 128, 256)}
SA_INPUT = {"sa2": 128, "sa3": 256, "sa4": 256}
FP_MLPS = {"fp1": (512, 256, 256), "fp2": (512, 256, 256)}


def synthetic():
    """situation3d_b200/synthetic.py as a stand-alone module (no package import)."""
    spec = importlib.util.spec_from_file_location("pn2_synthetic_scenes", os.path.join(_ROOT, "situation3d_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _layer(sd, prefix, cin, cout, g):
    w = torch.empty(cout, cin, 1, 1)
    torch.nn.init.kaiming_normal_(w, generator=g)
    sd[prefix + "conv.weight"] = w
    sd[prefix + "bn.bn.weight"] = 1.0 + 0.1 * torch.randn(cout, generator=g)
    sd[prefix + "bn.bn.bias"] = 0.1 * torch.randn(cout, generator=g)
    sd[prefix + "bn.bn.running_mean"] = 0.1 * torch.randn(cout, generator=g)
    sd[prefix + "bn.bn.running_var"] = torch.rand(cout, generator=g) + 0.5
    sd[prefix + "bn.bn.num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def backbone_state_dict(input_feature_dim=129, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, widths in SA_MLPS.items():
        cin = (input_feature_dim if name == "sa1" else SA_INPUT[name]) + 3
        for i, cout in enumerate(widths):
            _layer(sd, "%s.mlp_module.layer%d." % (name, i), cin, cout, g)
            cin = cout
    for name, widths in FP_MLPS.items():
        for i in range(len(widths) - 1):
            _layer(sd, "%s.mlp.layer%d." % (name, i), widths[i], widths[i + 1], g)
    return sd
