"""Drop-in mirror of the RF-Net / FDLNet descriptor `HardNetNeiMask`
(FDLNet-master/latency/rfnet/model/rf_des.py:11-116; the same class is model/des.py of the other latency experiments).

Same seven-conv body as HardNet without the Dropout, `input_norm` with eps 1e-8 and a plain `x / ||x||` head, plus the
neighbour-mask loss. eval() forward on CUDA tensors runs on the B200 kernels (the two epsilons go through
`hn_set_hardnet_eps`); train() forward and the loss are torch expressions.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .hardnet import HardNet
from .matching import distance_matrix_vector, pairwise_distances


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
