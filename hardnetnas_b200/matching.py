"""Drop-in mirror of the descriptor-matching part of FDLNet-master/utils/{math_utils,eval_utils}.py.

The reference materialises the full distance matrix and then takes `min` (eval_utils.py:113-114) or a full
row `sort` (:168-175). Here one fused pass (tensor-core GEMM + in-register shortlist, exact fp32 re-rank)
returns the nearest and second-nearest gallery row per query; nothing of size Nq x Ng is ever stored.
"""
from __future__ import annotations

import torch

from . import _ops


def assert_unit_norm(des, tol: float = 1e-3):
    """Precondition of the fused matching kernels (include/hardnet_b200.h, hn_match): rows are L2-normalised descriptors, the
    only input the FDLNet distance form `sqrt(2 - 2 a.b)` is defined for. The kernels do not check it (their fp16 operand
    packing and the re-rank margin assume |x| <= 1); call this once on descriptors of unknown origin - one device
    synchronisation, raises ValueError naming the first offending row."""
    n = des.float().norm(dim=1)
    bad = ((n - 1.0).abs() > tol) & (n > 0)        # all-zero rows (constant patches) are legal and match nothing
    if bool(bad.any()):
        i = int(bad.nonzero()[0])
        raise ValueError(f"descriptor row {i} has L2 norm {float(n[i]):.6f}: the matching kernels need unit-norm rows "
                         f"(|norm - 1| <= {tol})")
    return des


def distance_matrix_vector(anchor, positive):
    """Materialising API for small inputs (FDLNet-master/utils/math_utils.py:8-19)."""
    m = 2 - 2 * torch.mm(anchor, positive.t())
    return torch.sqrt(m.clamp(min=1e-8, max=4.0))


def pairwise_distances(x, y=None):
    """L2 distances between the rows of x and y (y = x when omitted), FDLNet-master/utils/math_utils.py:22-40:
    sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, min=1e-8)). Small coordinate sets only (keypoints); descriptors go through the fused
    matching kernel."""
    y = x if y is None else y
    d = (x * x).sum(1).view(-1, 1) + (y * y).sum(1).view(1, -1) - 2.0 * torch.mm(x, y.t())
    return torch.sqrt(d.clamp(min=1e-8))


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


def mutual_nn_ratio(des1, des2, threshold=0.7, return_pairs=True):
    """Mutual nearest neighbours AND the ratio test of BASELINE config 4 in one call:
    (pairs [M,2] int64 or None, mutual bool[Nq], ratio_label bool[Nq], Ia int64[Nq], Da, Db).
    (i, j) is mutual iff j = argmin_j D[i,:] and i = argmin_i D[:,j] - not in the reference; composed from its row / column
    minima (hardnet/Losses.py:105-108, eval_utils.py:24-32); ratio_label = Da / Db < threshold (eval_utils.py:168-175).
    ONE matching GEMM serves both directions (hn_match_mutual: the epilogue's per-cell maxima bound the column side, a small
    exact kernel verifies each column's best claimant). The masks come without any host synchronisation; compacting them into
    the `pairs` list needs its length on the host (return_pairs=False skips it)."""
    d1, d2, fwd, _, mutual = _ops.match_mutual(des1, des2)
    fwd = fwd.long()
    pairs = None
    if return_pairs:
        i = torch.arange(des1.size(0), device=fwd.device)
        pairs = torch.stack([i[mutual], fwd[mutual]], dim=1)
    return pairs, mutual, (d1 / d2).lt(threshold), fwd, d1, d2


def mutual_nn_ratio_two_pass(des1, des2, threshold=0.7, return_pairs=True):
    """The same result from two matching passes (queries x gallery, gallery x queries), operands packed once. Kept as the
    cross-check of the single-GEMM path (tests) and for A/B timing (bench.py)."""
    q16, g16 = _ops.pack_descriptors(des1), _ops.pack_descriptors(des2)
    d1, d2, fwd, _ = _ops.match_top2(des1, des2, q16=q16, g16=g16)
    _, _, bwd, _ = _ops.match_top2(des2, des1, q16=g16, g16=q16)
    fwd, bwd = fwd.long(), bwd.long()
    i = torch.arange(des1.size(0), device=fwd.device)
    mutual = bwd[fwd] == i
    pairs = torch.stack([i[mutual], fwd[mutual]], dim=1) if return_pairs else None
    return pairs, mutual, (d1 / d2).lt(threshold), fwd, d1, d2


def mutual_nearest_neighbors(des1, des2):
    """Mutual NN pairs [M,2]: (i, j) with j = argmin_j D[i,:] and i = argmin_i D[:,j] (see mutual_nn_ratio)."""
    return mutual_nn_ratio(des1, des2)[0]


# ---- match-score counters (FDLNet-master/utils/eval_utils.py:112-197) -----------------------------------------------
# The callers right behind the hot path in the reference's evaluation: nearest neighbour (or ratio test) on the fused
# kernel, then "is the matched keypoint within COO_THRSH pixels of the warped keypoint" and two counters. The reference
# builds an Nq x Nq coordinate distance matrix and takes its diagonal; the same expression is evaluated row-wise here.

def _keypoint_distance(kp1w, nn_kp2):
    """diag(pairwise_distances(kp1w[:, 1:3], nn_kp2[:, 1:3])) of math_utils.py:22-40: sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, 1e-8))."""
    x, y = kp1w[:, 1:3].float(), nn_kp2[:, 1:3].float()
    sq = (x * x).sum(1) + (y * y).sum(1) - 2.0 * (x * y).sum(1)
    return torch.sqrt(sq.clamp(min=1e-8))


def _score(predict_label, kp1w, nn_kp2, visible, coo_thrsh, predicted):
    correspondences = _keypoint_distance(kp1w, nn_kp2).le(coo_thrsh) & visible.bool()
    correct = (predict_label & correspondences).sum().item()
    return correct, max(predicted.sum().item(), 1)


def nearest_neighbor_match_score(des1, des2, kp1w, kp2, visible, COO_THRSH):
    """eval_utils.py:112-127 -> (correct_matches, predict_matches = max(#visible, 1))."""
    _, nn_idx = nearest_neighbor_match(des1, des2)
    vis = visible.bool()
    return _score(vis, kp1w, kp2.index_select(0, nn_idx), vis, COO_THRSH, vis)


def nearest_neighbor_threshold_match_score(des1, des2, kp1w, kp2, visible, DES_THRSH, COO_THRSH):
    """eval_utils.py:130-150 -> (correct_matches, predict_matches = max(#(nn_value < DES_THRSH and visible), 1))."""
    label, nn_idx = nearest_neighbor_threshold_match(des1, des2, DES_THRSH)
    predict = label & visible.bool()
    return _score(predict, kp1w, kp2.index_select(0, nn_idx), visible, COO_THRSH, predict)


def nearest_neighbor_distance_ratio_match_score(des1, des2, kp1w, kp2, visible, COO_THRSH, threshold=0.7):
    """eval_utils.py:178-197 -> (correct_matches, predict_matches = max(#(Da / Db < threshold and visible), 1))."""
    label, nn_kp2 = nearest_neighbor_distance_ratio_match(des1, des2, kp2, threshold)
    predict = label & visible.bool()
    return _score(predict, kp1w, nn_kp2, visible, COO_THRSH, predict)
