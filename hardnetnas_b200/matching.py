"""Drop-in mirror of the descriptor-matching part of FDLNet-master/utils/{math_utils,eval_utils}.py.

The reference materialises the full distance matrix and then takes `min` (eval_utils.py:113-114) or a full
row `sort` (:168-175). Here one fused pass (tensor-core GEMM + in-register shortlist, exact fp32 re-rank)
returns the nearest and second-nearest gallery row per query; nothing of size Nq x Ng is ever stored.
"""
from __future__ import annotations

import torch

from . import _ops


def distance_matrix_vector(anchor, positive):
    """Materialising API for small inputs (FDLNet-master/utils/math_utils.py:8-19)."""
    m = 2 - 2 * torch.mm(anchor, positive.t())
    return torch.sqrt(m.clamp(min=1e-8, max=4.0))


def nearest_neighbor_match(des1, des2):
    """`des_dist_matrix.min(dim=-1)` of eval_utils.py:113-114 -> (nn_value [Nq] fp32, nn_idx [Nq] int64)."""
    d1, _, i1, _ = _ops.match_top2(des1, des2)
    return d1, i1.long()


def nearest_neighbor_threshold_match(des1, des2, des_thrsh):
    """nn_value.lt(DES_THRSH) of eval_utils.py:133-135 -> (predict_label, nn_idx)."""
    d1, _, i1, _ = _ops.match_top2(des1, des2)
    return d1.lt(des_thrsh), i1.long()


def nearest_neighbor_distance_ratio_match(des1, des2, kp2, threshold):
    """eval_utils.py:168-175 -> (predict_label bool[Nq], nn_kp2 = kp2[Ia])."""
    d1, d2, i1, _ = _ops.match_top2(des1, des2)
    predict_label = (d1 / d2).lt(threshold)
    nn_kp2 = kp2.index_select(dim=0, index=i1.long().view(-1))
    return predict_label, nn_kp2


def match_top2(des1, des2):
    """(Da, Db, Ia, Ib) = sorted[:,0], sorted[:,1], indices[:,0], indices[:,1] of eval_utils.py:170-171."""
    d1, d2, i1, i2 = _ops.match_top2(des1, des2)
    return d1, d2, i1.long(), i2.long()


def mutual_nearest_neighbors(des1, des2):
    """Mutual NN pairs [M,2]: (i, j) with j = argmin_j D[i,:] and i = argmin_i D[:,j]. Not in the reference;
    composed from its row / column minima (hardnet/Losses.py:105-108, eval_utils.py:24-32)."""
    _, _, fwd, _ = _ops.match_top2(des1, des2)
    _, _, bwd, _ = _ops.match_top2(des2, des1)
    fwd, bwd = fwd.long(), bwd.long()
    i = torch.arange(des1.size(0), device=fwd.device)
    keep = bwd[fwd] == i
    return torch.stack([i[keep], fwd[keep]], dim=1)
