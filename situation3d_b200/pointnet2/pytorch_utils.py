"""Building blocks of the shared MLPs -- same public names, constructor arguments and
state-dict keys as the reference's lib/pointnet2/pytorch_utils.py:

    SharedMLP            pytorch_utils.py:11-36    layer{i} = Conv2d(1x1) [+ BN] [+ activation]
    BatchNorm1d/2d/3d    pytorch_utils.py:39-64    nn.Sequential holding one module named "bn"
    Conv1d/2d/3d         pytorch_utils.py:67-222   sub-modules "conv", "bn", "activation"
    FC                   pytorch_utils.py:225-260  sub-modules "fc", "bn", "activation"
    BNMomentumScheduler  pytorch_utils.py:271-296

so that a checkpoint written by the reference loads unchanged:
``...layer0.conv.weight``, ``...layer0.bn.bn.{weight,bias,running_mean,running_var,num_batches_tracked}``
(SURVEY.md A.7).  The fused CUDA path reads these parameters (folding BatchNorm in eval
mode); the modules themselves stay ordinary torch.nn modules so that training, .to(),
state_dict() and the unfused drop-in path behave exactly like the reference's.
"""
from typing import List, Tuple

import torch.nn as nn


def _norm_wrapper(norm_cls):
    class _Wrapped(nn.Sequential):
        def __init__(self, in_size: int, *, name: str = ""):
            super().__init__()
            norm = norm_cls(in_size)
            nn.init.constant_(norm.weight, 1.0)
            nn.init.constant_(norm.bias, 0)
            self.add_module(name + "bn", norm)

    return _Wrapped


class BatchNorm1d(_norm_wrapper(nn.BatchNorm1d)):
    pass


class BatchNorm2d(_norm_wrapper(nn.BatchNorm2d)):
    def __init__(self, in_size: int, name: str = ""):
        super().__init__(in_size, name=name)


class BatchNorm3d(_norm_wrapper(nn.BatchNorm3d)):
    def __init__(self, in_size: int, name: str = ""):
        super().__init__(in_size, name=name)


class _ConvBase(nn.Sequential):
    """conv (+ bn) (+ activation), or bn/activation first when ``preact``.  The conv carries a
    bias only when there is no BatchNorm (pytorch_utils.py:86)."""

    def __init__(self, in_size, out_size, kernel_size, stride, padding, activation, bn, init,
                 conv=None, batch_norm=None, bias=True, preact=False, name=""):
        super().__init__()
        use_bias = bias and not bn
        conv_unit = conv(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding,
                         bias=use_bias)
        init(conv_unit.weight)
        if use_bias:
            nn.init.constant_(conv_unit.bias, 0)
        norm_unit = batch_norm(in_size if preact else out_size) if bn else None

        tail = []
        if norm_unit is not None:
            tail.append((name + "bn", norm_unit))
        if activation is not None:
            tail.append((name + "activation", activation))
        ordered = tail + [(name + "conv", conv_unit)] if preact else [(name + "conv", conv_unit)] + tail
        for key, module in ordered:
            self.add_module(key, module)


class Conv1d(_ConvBase):
    def __init__(self, in_size: int, out_size: int, *, kernel_size: int = 1, stride: int = 1,
                 padding: int = 0, activation=nn.ReLU(inplace=True), bn: bool = False,
                 init=nn.init.kaiming_normal_, bias: bool = True, preact: bool = False, name: str = ""):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init,
                         conv=nn.Conv1d, batch_norm=BatchNorm1d, bias=bias, preact=preact, name=name)


class Conv2d(_ConvBase):
    def __init__(self, in_size: int, out_size: int, *, kernel_size: Tuple[int, int] = (1, 1),
                 stride: Tuple[int, int] = (1, 1), padding: Tuple[int, int] = (0, 0),
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_,
                 bias: bool = True, preact: bool = False, name: str = ""):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init,
                         conv=nn.Conv2d, batch_norm=BatchNorm2d, bias=bias, preact=preact, name=name)


class Conv3d(_ConvBase):
    def __init__(self, in_size: int, out_size: int, *, kernel_size: Tuple[int, int, int] = (1, 1, 1),
                 stride: Tuple[int, int, int] = (1, 1, 1), padding: Tuple[int, int, int] = (0, 0, 0),
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_,
                 bias: bool = True, preact: bool = False, name: str = ""):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init,
                         conv=nn.Conv3d, batch_norm=BatchNorm3d, bias=bias, preact=preact, name=name)


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d layers applied to (B, C, npoint, nsample).  With ``first`` and
    ``preact`` the very first layer has neither BatchNorm nor activation."""

    def __init__(self, args: List[int], *, bn: bool = False, activation=nn.ReLU(inplace=True),
                 preact: bool = False, first: bool = False, name: str = ""):
        super().__init__()
        for i, (cin, cout) in enumerate(zip(args[:-1], args[1:])):
            plain = first and preact and i == 0
            self.add_module(
                name + "layer{}".format(i),
                Conv2d(cin, cout, bn=bn and not plain, activation=None if plain else activation,
                       preact=preact))


class FC(nn.Sequential):
    def __init__(self, in_size: int, out_size: int, *, activation=nn.ReLU(inplace=True), bn: bool = False,
                 init=None, preact: bool = False, name: str = ""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if not bn:
            nn.init.constant_(fc.bias, 0)
        tail = []
        if bn:
            tail.append((name + "bn", BatchNorm1d(in_size if preact else out_size)))
        if activation is not None:
            tail.append((name + "activation", activation))
        ordered = tail + [(name + "fc", fc)] if preact else [(name + "fc", fc)] + tail
        for key, module in ordered:
            self.add_module(key, module)


def set_bn_momentum_default(bn_momentum):
    def fn(m):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            m.momentum = bn_momentum

    return fn


class BNMomentumScheduler(object):
    """Sets every BatchNorm's momentum to ``bn_lambda(epoch)`` (used by lib/solver.py:248-255)."""

    def __init__(self, model, bn_lambda, last_epoch=-1, setter=set_bn_momentum_default):
        if not isinstance(model, nn.Module):
            raise RuntimeError("Class '{}' is not a PyTorch nn Module".format(type(model).__name__))
        self.model = model
        self.setter = setter
        self.lmbd = bn_lambda
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
        self.last_epoch = epoch
        self.model.apply(self.setter(self.lmbd(epoch)))
