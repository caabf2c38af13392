"""CPU restatement of the distance / hardest-in-batch loss / matching path — test infrastructure only.

Follows hardnet/Losses.py:5-13,87-154, hardnetNAS/general_functions/Losses.py:27-51,
FDLNet-master/utils/math_utils.py:8-19, FDLNet-master/utils/eval_utils.py:112-175 and
hardnet/EvalMetrics.py:6-19 of the reference.
"""
from __future__ import annotations

import numpy as np
import torch


def distance_matrix_vector(anchor: torch.Tensor, positive: torch.Tensor) -> torch.Tensor:
    """hardnet/Losses.py:5-13 — sqrt(|a|^2 + |p|^2 - 2 a.p + 1e-6) with computed norms."""
    d1_sq = torch.sum(anchor * anchor, dim=1).unsqueeze(-1)
    d2_sq = torch.sum(positive * positive, dim=1).unsqueeze(-1)
    eps = 1e-6
    return torch.sqrt(d1_sq.repeat(1, positive.size(0)) + torch.t(d2_sq.repeat(1, anchor.size(0)))
                      - 2.0 * torch.mm(anchor, positive.t()) + eps)


def loss_hardnet_parts(anchor, positive, anchor_swap=False):
    """hardnet/Losses.py:95-108 — (pos, min_neg, row argmin, col argmin) of the 'min' batch reduce.

    The reference's `torch.eye(...).cuda()` (:96) is a device move only and is dropped here.
    """
    assert anchor.size() == positive.size(), "Input sizes between positive and negative must be equal."
    assert anchor.dim() == 2, "Inputd must be a 2D matrix."
    eps = 1e-8
    dist_matrix = distance_matrix_vector(anchor, positive) + eps
    eye = torch.eye(dist_matrix.size(1))
    pos1 = torch.diag(dist_matrix)
    dist_without_min_on_diag = dist_matrix + eye * 10
    mask = (dist_without_min_on_diag.ge(0.008).float() - 1.0) * (-1)
    mask = mask.type_as(dist_without_min_on_diag) * 10
    dist_without_min_on_diag = dist_without_min_on_diag + mask
    min_neg, row_arg = torch.min(dist_without_min_on_diag, 1)
    col_min, col_arg = torch.min(dist_without_min_on_diag, 0)
    if anchor_swap:
        min_neg = torch.min(min_neg, col_min)
    return pos1, min_neg, row_arg, col_min, col_arg


def loss_hardnet(anchor, positive, anchor_swap=False, margin=1.0) -> torch.Tensor:
    """hardnet/Losses.py:87-154 with batch_reduce='min', loss_type='triplet_margin'
    (hardnetNAS/general_functions/Losses.py:27-51 is anchor_swap=True)."""
    pos, min_neg, _, _, _ = loss_hardnet_parts(anchor, positive, anchor_swap)
    return torch.mean(torch.clamp(margin + pos - min_neg, min=0.0))


def distance_matrix_vector_fdl(anchor, positive) -> torch.Tensor:
    """FDLNet-master/utils/math_utils.py:8-19 — sqrt(clamp(2 - 2 a.p, 1e-8, 4))."""
    m = 2 - 2 * torch.mm(anchor, positive.t())
    return torch.sqrt(m.clamp(min=1e-8, max=4.0))


def nn_match(des1, des2, chunk: int = 2048):
    """D.min(dim=-1) of FDLNet-master/utils/eval_utils.py:113-114, evaluated in query-row chunks (the
    full matrix does not fit at 64k x 64k); row-wise identical to the reference."""
    vals, idxs = [], []
    for s in range(0, des1.size(0), chunk):
        d = distance_matrix_vector_fdl(des1[s:s + chunk], des2)
        v, i = d.min(dim=-1)
        vals.append(v)
        idxs.append(i)
    return torch.cat(vals), torch.cat(idxs)


def ratio_match(des1, des2, threshold: float = 0.7, chunk: int = 2048):
    """nearest_neighbor_distance_ratio_match, FDLNet-master/utils/eval_utils.py:168-175 (without the kp2
    gather): returns (predict_label, Ia, Da, Db). Uses topk(2, smallest) instead of a full sort — the two
    smallest values of a row are the same either way."""
    lab, ia, da, db = [], [], [], []
    for s in range(0, des1.size(0), chunk):
        d = distance_matrix_vector_fdl(des1[s:s + chunk], des2)
        v, i = torch.topk(d, 2, dim=-1, largest=False, sorted=True)
        a, b = v[:, 0], v[:, 1]
        lab.append((a / b).lt(threshold))
        ia.append(i[:, 0])
        da.append(a)
        db.append(b)
    return torch.cat(lab), torch.cat(ia), torch.cat(da), torch.cat(db)


def mutual_nn(des1, des2, chunk: int = 2048):
    """Mutual nearest neighbours. NOT in the reference; defined as the agreement of
    D.min(1) and D.min(0): i <-> j iff argmin_j D[i,:] == j and argmin_i D[:,j] == i. Returns [M,2]."""
    _, fwd = nn_match(des1, des2, chunk)
    _, bwd = nn_match(des2, des1, chunk)
    i = torch.arange(des1.size(0))
    keep = bwd[fwd] == i
    return torch.stack([i[keep], fwd[keep]], dim=1)


def error_rate_at_95_recall(labels: np.ndarray, scores: np.ndarray) -> float:
    """hardnet/EvalMetrics.py:6-19."""
    distances = 1.0 / (scores + 1e-8)
    recall_point = 0.95
    labels = labels[np.argsort(distances)]
    threshold_index = np.argmax(np.cumsum(labels) >= recall_point * np.sum(labels))
    fp = np.sum(labels[:threshold_index] == 0)
    tn = np.sum(labels[threshold_index:] == 0)
    return float(fp) / float(fp + tn)
