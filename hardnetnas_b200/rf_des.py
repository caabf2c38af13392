"""Drop-in mirror of the RF-Net / FDLNet descriptor `HardNetNeiMask`
(FDLNet-master/latency/rfnet/model/rf_des.py:11-116; the same class is model/des.py of the other latency experiments).

Same seven-conv body as HardNet without the Dropout, `input_norm` with eps 1e-8 and a plain `x / ||x||` head, plus the
neighbour-mask loss. eval() forward on CUDA tensors runs on the B200 kernels (the two epsilons go through
`hn_set_hardnet_eps`); the loss of CUDA descriptors runs on the fused distance kernel with the keypoint masks as an input
of its epilogue (`hn_dist_min_ex`, sparse autograd backward); train() forward and the loss of CPU tensors are torch expressions.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, _ops
from .hardnet import HardNet
from .matching import distance_matrix_vector, pairwise_distances


class _FusedNeighbourMaskLoss(torch.autograd.Function):
    """HardNetNeiMask.loss on the fused distance kernel (hn_dist_min_ex, HN_FLAG_NEI_MASK): the N x N distance matrix, the
    diagonal and the two keypoint-neighbourhood masks and the row / column minima happen in the GEMM epilogue. Backward:
    gradients flow through pos[i] and the one selected negative of each active row (the +10 offsets are constants), rebuilt
    from the arg-indices the kernel returns; d = sqrt(clamp(2 - 2 a.p, 1e-8, 4)) has gradient -p / d (0 where clamped)."""

    @staticmethod
    def forward(ctx, anchor, positive, a_xy, p_xy, margin, c):
        res = _ops.dist_min(anchor, positive, _lib.HN_FORM_FDL, loss_mask=False, swap=True, a_xy=a_xy, p_xy=p_xy, nei_c=c)
        pos, row_min, col_min = res["pos"], res["row_min"], res["col_min"]
        hardest = torch.minimum(row_min, col_min)
        ctx.save_for_backward(anchor, positive, pos, hardest, row_min, col_min, res["row_arg"], res["col_arg"])
        ctx.margin = margin
        return torch.clamp(margin + pos - hardest, min=0.0).mean()

    @staticmethod
    def backward(ctx, grad_out):
        anchor, positive, pos, hardest, row_min, col_min, row_arg, col_arg = ctx.saved_tensors
        a, p = anchor.detach().float(), positive.detach().float()
        n = a.size(0)
        active = ((ctx.margin + pos - hardest) > 0).float() * (grad_out / n)

        def pair_grad(x, y):                        # d/dx of sqrt(clamp(2 - 2 x.y, 1e-8, 4)) = -y / d inside the clamp
            m = 2.0 - 2.0 * (x * y).sum(1)
            inside = ((m > 1e-8) & (m < 4.0)).float()
            d = torch.sqrt(m.clamp(min=1e-8, max=4.0))
            return (inside / d).unsqueeze(1)

        ga, gp = torch.zeros_like(a), torch.zeros_like(p)
        w = active.unsqueeze(1) * pair_grad(a, p)   # + pos[i]
        ga -= w * p
        gp -= w * a
        use_col = col_min < row_min
        j = row_arg.long()                          # - masked[i, j*] for rows whose hardest negative sits in their row
        w = ((~use_col).float() * active).unsqueeze(1) * pair_grad(a, p[j])
        ga += w * p[j]
        gp.index_add_(0, j, w * a)
        k = col_arg.long()                          # - masked[k*, i] for rows whose hardest negative sits in their column
        w = (use_col.float() * active).unsqueeze(1) * pair_grad(a[k], p)
        ga.index_add_(0, k, w * p)
        gp += w * a[k]
        return ga.to(anchor.dtype), gp.to(positive.dtype), None, None, None, None


class HardNetNeiMask(HardNet):
    INPUT_NORM_EPS = 1e-8    # rf_des.py:43
    L2_EPS = 0.0             # rf_des.py:54: x / torch.norm(x, p=2, dim=-1, keepdim=True)

    def __init__(self, MARGIN, C, act_dtype: str = "fp16", chunk_patches: int = 0, head_rows: int = 0):
        super().__init__(act_dtype=act_dtype, chunk_patches=chunk_patches, head_rows=head_rows)
        self.MARGIN = MARGIN
        self.C = C
        # the reference's Sequential has no Dropout: indices 0..19 (state_dict keys features.{0,1,3,4,...,18,19})
        self.features = nn.Sequential(*[m for m in self.features if not isinstance(m, nn.Dropout)])

    def forward(self, input, out_dtype: torch.dtype = torch.float32, out: torch.Tensor | None = None):
        if self.training:
            x = self.features(self.input_norm(input))
            x = x.view(x.size(0), -1)
            return x / torch.norm(x, p=2, dim=-1, keepdim=True)
        return self._forward_b200(input, out_dtype, out)

    def loss(self, anchor, positive, anchor_kp, positive_kp):
        """Hardest-in-batch margin loss where, besides the diagonal, every pair whose keypoints lie closer than C pixels
        (in either image) is excluded from the negatives (rf_des.py:57-96). With C = 0 it is the plain hard loss."""
        assert anchor.size() == positive.size()
        assert anchor.dim() == 2
        if anchor.is_cuda and anchor.size(1) == 128:
            return _FusedNeighbourMaskLoss.apply(anchor, positive, anchor_kp[:, 1:3].to(torch.float), positive_kp[:, 1:3].to(torch.float),
                                                 float(self.MARGIN), float(self.C))
        d = distance_matrix_vector(anchor, positive)
        pos = d.diag()
        masked = d + torch.eye(d.size(1), device=d.device, dtype=d.dtype) * 10
        for kp in (anchor_kp, positive_kp):
            near = pairwise_distances(kp[:, 1:3].to(torch.float)).lt(self.C)
            masked = masked + near.to(torch.float) * 10
        hardest = torch.min(masked.min(dim=1)[0], masked.min(dim=0)[0])
        return torch.clamp(self.MARGIN + pos - hardest, min=0.0).mean()

    @staticmethod
    def weights_init(m):
        if isinstance(m, nn.Conv2d):
            nn.init.orthogonal_(m.weight.data, gain=0.6)
