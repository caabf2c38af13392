"""CPU restatement of HardNet.forward (eval mode) — test infrastructure only (see oracle/__init__.py).

Follows hardnet/HardNet.py:275-315 and hardnet/Utils.py:15-22 of the reference, fp32 torch CPU ops.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .synth import BN_EPS, HARDNET_STAGES


def input_norm(x: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """hardnet/HardNet.py:306-310 — per-patch (x - mean) / (unbiased std + 1e-7); HardNetNeiMask uses 1e-8
    (FDLNet-master/latency/rfnet/model/rf_des.py:41-49)."""
    flat = x.reshape(x.size(0), -1)
    mp = flat.mean(dim=1)
    sp = flat.std(dim=1) + eps
    return (x - mp.view(-1, 1, 1, 1)) / sp.view(-1, 1, 1, 1)


def l2norm(x: torch.Tensor, eps: float = 1e-10) -> torch.Tensor:
    """hardnet/Utils.py:15-22 — x / sqrt(sum(x*x, 1) + 1e-10)."""
    norm = torch.sqrt(torch.sum(x * x, dim=1) + eps)
    return x / norm.unsqueeze(-1)


def hardnet_stages(x, weights, bn_means, bn_vars, upto: int = 7, norm_eps: float = 1e-7):
    """Activations after each conv+BN(+ReLU) stage, NCHW fp32 (hardnet/HardNet.py:280-302).

    BatchNorm is eval-mode, affine=False: (y - running_mean) / sqrt(running_var + 1e-5). Dropout(0.3)
    before the last conv is the identity in eval mode.
    """
    acts = []
    h = input_norm(x, norm_eps)
    for i, (cin, cout, k, stride, pad) in enumerate(HARDNET_STAGES[:upto]):
        h = F.conv2d(h, weights[i], None, stride=stride, padding=pad)
        h = (h - bn_means[i].view(1, -1, 1, 1)) / torch.sqrt(bn_vars[i].view(1, -1, 1, 1) + BN_EPS)
        if i < 6:
            h = F.relu(h)
        acts.append(h)
    return acts


def hardnet_forward(x, weights, bn_means, bn_vars, norm_eps: float = 1e-7, l2_eps: float = 1e-10) -> torch.Tensor:
    """hardnet/HardNet.py:312-315 — descriptors [B,128], unit rows (zero row for an all-zero feature).
    norm_eps = 1e-8, l2_eps = 0 restate HardNetNeiMask.forward (FDLNet-master/latency/rfnet/model/rf_des.py:41-55)."""
    with torch.no_grad():
        feats = hardnet_stages(x, weights, bn_means, bn_vars, norm_eps=norm_eps)[-1]
        return l2norm(feats.reshape(feats.size(0), -1), l2_eps)
