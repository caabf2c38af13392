"""Building blocks of the NAS search space: the ops reachable from CANDIDATE_BLOCKS
(hardnetNAS/supernet_functions/lookup_table_builder.py:18-20) through PRIMITIVES
(hardnetNAS/fbnet_building_blocks/fbnet_builder.py:36-155).

Class names, constructor contracts, submodule names (`pw`, `dw`, `pwl`, `se4`, `conv`, `bn`, `op`) and parameter
creation order follow the reference, so state_dicts are interchangeable and `torch.manual_seed(s)` gives the same
initial weights. These modules are plain torch (definition / training surface); the accelerated eval forward
lives in `SampledDescriptorNet`, which compiles a stack of them into an op list for the C ABI.

Not provided (not in CANDIDATE_BLOCKS): Shift / ShiftBlock5x5 / CascadeConv3x3 / Upsample,
k7 and `cdw` variants, group-norm / frozen-BN flavours.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class Flatten(nn.Module):
    """[N, C, H, W] -> [N, C*H*W] (fbnet_builder.py:193-199)."""

    def forward(self, x):
        return x.view(-1, int(torch.tensor(x.shape[1:]).prod().item()))


class ChannelShuffle(nn.Module):
    """[N,C,H,W] -> [N,g,C/g,H,W] -> [N,C/g,g,H,W] -> [N,C,H,W] (fbnet_builder.py:332-349)."""

    def __init__(self, groups):
        super().__init__()
        self.groups = groups

    def forward(self, x):
        n, c, h, w = x.size()
        g = self.groups
        assert c % g == 0, "Incompatible group size {} for input channel {}".format(g, c)
        return x.view(n, g, c // g, h, w).permute(0, 2, 1, 3, 4).contiguous().view(n, c, h, w)

    def source_channels(self, c: int) -> torch.Tensor:
        """perm with shuffled[:, j] == x[:, perm[j]] — used to fold the shuffle into the producing conv."""
        g = self.groups
        j = torch.arange(c)
        return (j % g) * (c // g) + j // g


class ConvBNRelu(nn.Sequential):
    """conv (kaiming-normal fan_out) [+ BatchNorm2d] [+ ReLU] (fbnet_builder.py:352-404)."""

    def __init__(self, input_depth, output_depth, kernel, stride, pad, no_bias, use_relu, bn_type, group=1, *args, **kwargs):
        super().__init__()
        assert use_relu in ["relu", None]
        assert bn_type in ["bn", None], "only BatchNorm2d ('bn') or no norm is supported on this path"
        assert stride in [1, 2, 4]
        op = nn.Conv2d(input_depth, output_depth, kernel_size=kernel, stride=stride, padding=pad, bias=not no_bias,
                       groups=group, *args, **kwargs)
        nn.init.kaiming_normal_(op.weight, mode="fan_out", nonlinearity="relu")
        if op.bias is not None:
            nn.init.constant_(op.bias, 0.0)
        self.add_module("conv", op)
        if bn_type == "bn":
            self.add_module("bn", nn.BatchNorm2d(output_depth))
        if use_relu == "relu":
            self.add_module("relu", nn.ReLU(inplace=True))


class SEModule(nn.Module):
    """x * sigmoid(fc2(relu(fc1(global_avg_pool(x))))) (fbnet_builder.py:407-421)."""
    reduction = 4

    def __init__(self, C):
        super().__init__()
        mid = max(C // self.reduction, 8)
        conv1 = nn.Conv2d(C, mid, 1, 1, 0)
        conv2 = nn.Conv2d(mid, C, 1, 1, 0)
        self.op = nn.Sequential(nn.AdaptiveAvgPool2d(1), conv1, nn.ReLU(inplace=True), conv2, nn.Sigmoid())

    def forward(self, x):
        return x * self.op(x)


class Identity(nn.Module):
    """The `skip` candidate: nothing, MaxPool(3,2,1), a 1x1 ConvBNRelu, or both (fbnet_builder.py:202-228)."""

    def __init__(self, C_in, C_out, stride):
        super().__init__()
        self.output_depth = C_out
        pool = [nn.MaxPool2d(kernel_size=3, stride=2, padding=1)] if stride != 1 else []
        proj = ([ConvBNRelu(C_in, C_out, kernel=1, stride=1, pad=0, no_bias=1, use_relu="relu", bn_type="bn")]
                if C_in != C_out else [])
        if not pool and not proj:
            self.conv = None
        elif not pool:
            self.conv = proj[0]
        else:
            self.conv = nn.Sequential(*pool, *proj)

    def forward(self, x):
        return self.conv(x) if self.conv is not None else x


class IRFBlock(nn.Module):
    """Inverted residual: 1x1 (grouped) conv+BN+ReLU -> [channel shuffle] -> depthwise kxk (stride) +BN+ReLU ->
    1x1 (grouped) conv+BN -> [+x if stride 1 and C_in == C_out] -> [SE] (fbnet_builder.py:455-570)."""

    def __init__(self, input_depth, output_depth, expansion, stride, bn_type="bn", kernel=3, width_divisor=1,
                 shuffle_type=None, pw_group=1, se=False, cdw=False, dw_skip_bn=False, dw_skip_relu=False):
        super().__init__()
        assert kernel in [3, 5], "only k3 / k5 depthwise kernels are in the searched candidate set"
        assert not cdw and width_divisor == 1 and stride in (1, 2)
        self.use_res_connect = stride == 1 and input_depth == output_depth
        self.output_depth = output_depth
        mid_depth = int(input_depth * expansion)
        self.pw = ConvBNRelu(input_depth, mid_depth, kernel=1, stride=1, pad=0, no_bias=1, use_relu="relu",
                             bn_type=bn_type, group=pw_group)
        self.upscale = None
        self.dw = ConvBNRelu(mid_depth, mid_depth, kernel=kernel, stride=stride, pad=kernel // 2, group=mid_depth,
                             no_bias=1, use_relu="relu" if not dw_skip_relu else None,
                             bn_type=bn_type if not dw_skip_bn else None)
        self.pwl = ConvBNRelu(mid_depth, output_depth, kernel=1, stride=1, pad=0, no_bias=1, use_relu=None,
                              bn_type=bn_type, group=pw_group)
        self.shuffle_type = shuffle_type
        if shuffle_type is not None:
            self.shuffle = ChannelShuffle(pw_group)
        self.se4 = SEModule(output_depth) if se else nn.Sequential()

    def forward(self, x):
        y = self.pw(x)
        if self.shuffle_type == "mid":
            y = self.shuffle(y)
        y = self.pwl(self.dw(y))
        if self.use_res_connect:
            y = y + x
        return self.se4(y)


def _irf(expansion, kernel, **fixed):
    def make(C_in, C_out, _expansion, stride, **kwargs):
        return IRFBlock(C_in, C_out, expansion, stride, kernel=kernel, **fixed, **kwargs)
    return make


# name -> constructor(C_in, C_out, expansion(ignored), stride); the expansion is embedded in the op name
PRIMITIVES = {"skip": lambda C_in, C_out, expansion, stride, **kwargs: Identity(C_in, C_out, stride)}
for _k in (3, 5):
    for _e in (1, 3, 6):
        PRIMITIVES[f"ir_k{_k}_e{_e}"] = _irf(_e, _k)
        PRIMITIVES[f"ir_k{_k}_e{_e}_se"] = _irf(_e, _k, se=True)
    PRIMITIVES[f"ir_k{_k}_s4"] = _irf(4, _k, shuffle_type="mid", pw_group=4)
    PRIMITIVES[f"ir_k{_k}_s4_se"] = _irf(4, _k, shuffle_type="mid", pw_group=4, se=True)
    PRIMITIVES[f"ir_k{_k}_s2"] = _irf(1, _k, shuffle_type="mid", pw_group=2)
    PRIMITIVES[f"ir_k{_k}_s2_se"] = _irf(1, _k, shuffle_type="mid", pw_group=2, se=True)
