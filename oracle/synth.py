"""Seeded synthetic inputs shared by the oracle, the parity tests and bench.py.

Everything is generated on the CPU with explicit torch.Generator seeds so that the golden vectors made
in the build container can be regenerated bit-identically on the GPU box.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# HardNet stage shapes: (C_in, C_out, kernel, stride, padding), hardnet/HardNet.py:280-302
HARDNET_STAGES = [
    (1, 32, 3, 1, 1),
    (32, 32, 3, 1, 1),
    (32, 64, 3, 2, 1),
    (64, 64, 3, 1, 1),
    (64, 128, 3, 2, 1),
    (128, 128, 3, 1, 1),
    (128, 128, 8, 1, 0),
]
CONV_IDX = [0, 3, 6, 9, 12, 15, 19]  # nn.Sequential indices of the convs
BN_IDX = [1, 4, 7, 10, 13, 16, 20]   # ... and of the BatchNorms
BN_EPS = 1e-5


def make_patches(n: int, seed: int = 1234, edge_cases: bool = True) -> torch.Tensor:
    """[n,1,32,32] fp32 in [0,1): uniform noise smoothed 5x5 so the patches have structure.

    With edge_cases the last two patches are a constant patch (input_norm -> exactly 0) and a patch with
    a tiny standard deviation.
    """
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 1, 32, 32, generator=g)
    x = F.avg_pool2d(x, 5, 1, 2)
    if edge_cases and n >= 4:
        x[-1] = 0.5
        x[-2] = 0.25 + 1e-4 * x[-2]
    return x.contiguous()


def make_positives(anchors: torch.Tensor, sigma: float = 0.1, seed: int = 7) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (anchors + sigma * torch.randn(anchors.shape, generator=g)).contiguous()


def randomize_bn_stats(state_dict: dict, seed: int = 3) -> dict:
    """running_mean ~ 0.1 N(0,1), running_var ~ U(0.5,1.5): without this eval-mode BN is a constant scale
    and the fold would go untested."""
    g = torch.Generator().manual_seed(seed)
    out = dict(state_dict)
    for k in sorted(out.keys()):
        if k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(out[k].shape, generator=g)
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(out[k].shape, generator=g)
    return out


def hardnet_weights_from_seed(seed: int = 0, bn_seed: int | None = 3):
    """Weights exactly as the reference draws them: torch.manual_seed(seed); HardNet() applies
    nn.init.orthogonal_(gain=0.6) to each conv in module order (hardnet/HardNet.py:303,317-324).

    Returns (list of 7 conv weights, list of 7 running_mean, list of 7 running_var).
    """
    torch.manual_seed(seed)
    ws = []
    for cin, cout, k, _, _ in HARDNET_STAGES:
        conv = torch.nn.Conv2d(cin, cout, kernel_size=k, bias=False)
        ws.append(conv)
    # the reference constructs all modules first (default init consumes RNG), then applies weights_init
    for conv in ws:
        torch.nn.init.orthogonal_(conv.weight.data, gain=0.6)
    w = [c.weight.data.clone() for c in ws]
    means = [torch.zeros(s[1]) for s in HARDNET_STAGES]
    vars_ = [torch.ones(s[1]) for s in HARDNET_STAGES]
    if bn_seed is not None:
        sd = {}
        for i, bi in enumerate(BN_IDX):
            sd[f"features.{bi}.running_mean"] = means[i]
            sd[f"features.{bi}.running_var"] = vars_[i]
        sd = randomize_bn_stats(sd, bn_seed)
        means = [sd[f"features.{bi}.running_mean"] for bi in BN_IDX]
        vars_ = [sd[f"features.{bi}.running_var"] for bi in BN_IDX]
    return w, means, vars_


def unit_vectors(n: int, d: int = 128, seed: int = 11) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(n, d, generator=g)
    return (v / v.norm(dim=1, keepdim=True)).contiguous()


def make_match_set(nq: int, ng: int, seed: int = 11, sigma: float = 0.04, match_frac: float = 0.8):
    """Gallery of unit vectors; 80 % of the queries are noisy copies of distinct gallery rows (true match
    at d ~ 0.43 vs ~1.41 for distractors), 20 % are fresh unit vectors (no match). Returns
    (queries, gallery, true_index) with true_index = -1 for distractor queries."""
    gal = unit_vectors(ng, 128, seed)
    g = torch.Generator().manual_seed(seed + 1)
    n_match = min(int(nq * match_frac), ng)
    perm = torch.randperm(ng, generator=g)[:n_match]
    q = torch.empty(nq, 128)
    q[:n_match] = gal[perm] + sigma * torch.randn(n_match, 128, generator=g)
    q[n_match:] = torch.randn(nq - n_match, 128, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    truth = torch.full((nq,), -1, dtype=torch.long)
    truth[:n_match] = perm
    shuffle = torch.randperm(nq, generator=g)
    return q[shuffle].contiguous(), gal, truth[shuffle].contiguous()


def weights_fingerprint(ws) -> np.ndarray:
    """Small fingerprint of a weight list (sum, abs-sum and first 8 values per tensor) stored with the
    golden vectors so a test can prove the regenerated weights are the ones the reference produced."""
    rows = []
    for w in ws:
        f = w.double().flatten()
        rows.append(np.concatenate([[f.sum().item(), f.abs().sum().item()], f[:8].numpy()]))
    return np.stack(rows)


def randomize_nas_state(state_dict: dict, seed: int = 4) -> dict:
    """Randomise every BatchNorm of a NAS net (running stats and, where present, the affine parameters):
    running_mean ~ 0.1 N, running_var ~ U(0.5,1.5), weight ~ U(0.5,1.5), bias ~ 0.1 N. Keys are visited in sorted
    order so that two state_dicts with the same keys receive the same values."""
    g = torch.Generator().manual_seed(seed)
    out = dict(state_dict)
    bn_prefixes = sorted(k[: -len("running_mean")] for k in out if k.endswith("running_mean"))
    for pre in bn_prefixes:
        n = out[pre + "running_mean"].numel()
        out[pre + "running_mean"] = 0.1 * torch.randn(n, generator=g)
        out[pre + "running_var"] = 0.5 + torch.rand(n, generator=g)
        if pre + "weight" in out:
            out[pre + "weight"] = 0.5 + torch.rand(n, generator=g)
            out[pre + "bias"] = 0.1 * torch.randn(n, generator=g)
    return out
