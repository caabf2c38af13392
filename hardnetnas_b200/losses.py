"""Drop-in mirror of hardnet/Losses.py (and hardnetNAS/general_functions/Losses.py) for the hot path.

`loss_HardNet(..., batch_reduce='min', loss_type='triplet_margin')` — the configuration the reference trains
with (hardnet/HardNet.py:408-413) — runs as one fused B200 pass: tensor-core distance GEMM whose epilogue
applies the eps / diagonal / duplicate masks and the row (and column) minima. The N x N matrix is never
materialised. The other reduce / loss modes are elementwise variants outside the accelerated path and are
evaluated with the same torch expression sequence as the reference.
"""
from __future__ import annotations

import sys

import torch

from . import _lib, _ops


def distance_matrix_vector(anchor, positive):
    """Materialising API, kept for small N (hardnet/Losses.py:5-13). The fused paths never call it."""
    d1_sq = torch.sum(anchor * anchor, dim=1).unsqueeze(-1)
    d2_sq = torch.sum(positive * positive, dim=1).unsqueeze(-1)
    eps = 1e-6
    return torch.sqrt((d1_sq + d2_sq.t()) - 2.0 * torch.mm(anchor, positive.t()) + eps)


def _masked_matrix(anchor, positive):
    eps = 1e-8
    dist_matrix = distance_matrix_vector(anchor, positive) + eps
    eye = torch.eye(dist_matrix.size(1), device=dist_matrix.device, dtype=dist_matrix.dtype)
    pos1 = torch.diag(dist_matrix)
    d = dist_matrix + eye * 10
    mask = (d.ge(0.008).float() - 1.0) * (-1)
    d = d + mask.type_as(d) * 10
    return pos1, d


class _FusedHardestInBatch(torch.autograd.Function):
    """Forward: fused kernel. Backward: gradients only flow through pos[i] and the selected negative of each
    row (<= 2 non-zeros per row), rebuilt from the arg-indices the fused kernel returns."""

    @staticmethod
    def forward(ctx, anchor, positive, margin, anchor_swap):
        res = _ops.dist_min(anchor, positive, _lib.HN_FORM_HARDNET, loss_mask=True, swap=anchor_swap)
        pos, row_min = res["pos"], res["row_min"]
        min_neg = torch.minimum(row_min, res["col_min"]) if anchor_swap else row_min
        per_row = torch.clamp(margin + pos - min_neg, min=0.0)
        ctx.save_for_backward(anchor, positive, pos, min_neg, row_min, res["row_arg"],
                              res["col_min"] if anchor_swap else row_min, res["col_arg"] if anchor_swap else res["row_arg"])
        ctx.margin, ctx.swap = margin, anchor_swap
        return per_row.mean()

    @staticmethod
    def backward(ctx, grad_out):
        anchor, positive, pos, min_neg, row_min, row_arg, col_min, col_arg = ctx.saved_tensors
        n = anchor.size(0)
        a, p = anchor.detach().float(), positive.detach().float()
        active = ((ctx.margin + pos - min_neg) > 0).float() * (grad_out / n)
        ga = torch.zeros_like(a)
        gp = torch.zeros_like(p)
        # d/da of sqrt(|a-p|^2 + 1e-6) = (a - p) / d ; pos already carries the +1e-8
        diff = a - p
        w = (active / (pos - 1e-8)).unsqueeze(1)
        ga += w * diff
        gp -= w * diff
        idx = torch.arange(n, device=a.device)
        use_col = (col_min < row_min) if ctx.swap else torch.zeros(n, dtype=torch.bool, device=a.device)
        # negative taken from the row: pair (a_i, p_j*), masked values (+10) have the same gradient as d
        j = row_arg.long()
        sel = (~use_col).float() * active
        dneg = (a - p[j])
        dn = torch.sqrt((dneg * dneg).sum(1) + 1e-6).unsqueeze(1)
        g = (sel.unsqueeze(1) * dneg / dn)
        ga -= g
        gp.index_add_(0, j, g)
        if ctx.swap:
            # negative taken from the column: pair (a_k*, p_i)
            k = col_arg.long()
            selc = use_col.float() * active
            dneg = (a[k] - p)
            dn = torch.sqrt((dneg * dneg).sum(1) + 1e-6).unsqueeze(1)
            g = (selc.unsqueeze(1) * dneg / dn)
            ga.index_add_(0, k, -g)
            gp += g
        del idx
        return ga.to(anchor.dtype), gp.to(positive.dtype), None, None


def loss_HardNet(anchor, positive, anchor_swap=False, anchor_ave=False, margin=1.0, batch_reduce='min',
                 loss_type="triplet_margin"):
    """HardNet margin loss (hardnet/Losses.py:87-154): positive distance vs closest in-batch negative."""
    assert anchor.size() == positive.size(), "Input sizes between positive and negative must be equal."
    assert anchor.dim() == 2, "Inputd must be a 2D matrix."
    if batch_reduce == 'min' and loss_type == "triplet_margin":
        if torch.is_grad_enabled() and (anchor.requires_grad or positive.requires_grad):
            return _FusedHardestInBatch.apply(anchor, positive, float(margin), bool(anchor_swap))
        return _ops.loss_hardnet(anchor, positive, float(margin), bool(anchor_swap))
    # ---- modes outside the accelerated path (hardnet/Losses.py:109-152): which negatives enter, then the shared reduction ----
    pos, d = _masked_matrix(anchor, positive)
    n = anchor.size(0)
    if batch_reduce == 'min':
        neg, neg_t = d.min(dim=1)[0], d.min(dim=0)[0]
    elif batch_reduce == 'average':
        pos = pos.repeat(n)                                  # every (row, column) pair is a term
        neg, neg_t = d.reshape(-1), d.t().reshape(-1)
    elif batch_reduce == 'random':
        idxs = torch.randperm(n, device=anchor.device).view(-1, 1)
        neg, neg_t = d.gather(1, idxs).view(-1), d.t().gather(1, idxs).view(-1)
    else:
        print('Unknown batch reduce mode. Try min, average or random')
        sys.exit(1)
    min_neg = torch.min(neg, neg_t) if anchor_swap else neg
    return torch.mean(_reduce_triplet(pos, min_neg, margin, loss_type))


def loss_HardNet_nas(anchor, positive, margin=1.0):
    """hardnetNAS/general_functions/Losses.py:27-51 — same loss with anchor swap always on."""
    return loss_HardNet(anchor, positive, anchor_swap=True, margin=margin)


# ---- the other loss helpers of hardnet/Losses.py -----------------------------------------------------------------
# O(N * D) or small-N training utilities next to the hot path; hardnet/HardNet.py:36 imports them by name, so a module
# swap needs them. Plain torch expressions on whatever device the inputs live on (no hard-coded .cuda()).

def distance_vectors_pairwise(anchor, positive, negative=None):
    """Row-wise L2 distances d(a_i, p_i) [and d(a_i, n_i), d(p_i, n_i)] with the reference's +1e-8 under the root
    (hardnet/Losses.py:15-27)."""
    eps = 1e-8
    sq_a = (anchor * anchor).sum(dim=1)
    sq_p = (positive * positive).sum(dim=1)

    def dist(sq_x, sq_y, x, y):
        return torch.sqrt(sq_x + sq_y - 2 * (x * y).sum(dim=1) + eps)

    d_ap = dist(sq_a, sq_p, anchor, positive)
    if negative is None:
        return d_ap
    sq_n = (negative * negative).sum(dim=1)
    return d_ap, dist(sq_a, sq_n, anchor, negative), dist(sq_p, sq_n, positive, negative)


def _reduce_triplet(pos, min_neg, margin, loss_type, eps=1e-8):
    if loss_type == "triplet_margin":
        return torch.clamp(margin + pos - min_neg, min=0.0)
    if loss_type == "softmax":
        e_pos = torch.exp(2.0 - pos)
        return -torch.log(e_pos / (e_pos + torch.exp(2.0 - min_neg) + eps))
    if loss_type == "contrastive":
        return torch.clamp(margin - min_neg, min=0.0) + pos
    print('Unknown loss type. Try triplet_margin, softmax or contrastive')
    sys.exit(1)


def loss_random_sampling(anchor, positive, negative, anchor_swap=False, margin=1.0, loss_type="triplet_margin"):
    """Triplet-style loss with given (random) negatives instead of in-batch mining (hardnet/Losses.py:29-55)."""
    assert anchor.size() == positive.size(), "Input sizes between positive and negative must be equal."
    assert anchor.size() == negative.size(), "Input sizes between positive and negative must be equal."
    assert anchor.dim() == 2, "Inputd must be a 2D matrix."
    pos, d_an, d_pn = distance_vectors_pairwise(anchor, positive, negative)
    min_neg = torch.min(d_an, d_pn) if anchor_swap else d_an
    return torch.mean(_reduce_triplet(pos, min_neg, margin, loss_type))


def loss_L2Net(anchor, positive, anchor_swap=False, margin=1.0, loss_type="triplet_margin"):
    """L2Net sampling: the whole batch is the negative set; only the softmax form exists (hardnet/Losses.py:57-85).
    The reference also builds the diagonal / duplicate masks here but never uses them in this branch."""
    assert anchor.size() == positive.size(), "Input sizes between positive and negative must be equal."
    assert anchor.dim() == 2, "Inputd must be a 2D matrix."
    if loss_type != 'softmax':
        print('Only softmax loss works with L2Net sampling')
        sys.exit(1)
    eps = 1e-8
    d = distance_matrix_vector(anchor, positive)
    e = torch.exp(2.0 - d)
    e_pos = torch.exp(2.0 - torch.diag(d))
    loss = -torch.log(e_pos / (e.sum(dim=1) + eps))
    if anchor_swap:
        loss = loss - torch.log(e_pos / (e.sum(dim=0) + eps))
    return torch.mean(loss)


def global_orthogonal_regularization(anchor, negative):
    """GOR (hardnet/Losses.py:156-162): (mean a.n)^2 + max(mean (a.n)^2 - 1/d, 0)."""
    dots = (anchor * negative).sum(dim=1)
    return dots.mean() ** 2 + torch.clamp((dots * dots).mean() - 1.0 / anchor.size(1), min=0.0)
