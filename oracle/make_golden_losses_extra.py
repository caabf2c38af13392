"""Golden values for the loss helpers next to loss_HardNet in hardnet/Losses.py (distance_vectors_pairwise :15-27,
loss_random_sampling :29-55, loss_L2Net :57-85, global_orthogonal_regularization :156-162), produced by importing the
UNMODIFIED reference module in this container. TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_losses_extra.py      (needs /root/reference; writes tests/golden/losses_extra.npz)
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import synth  # noqa: E402

REF = Path("/root/reference")


def inputs():
    a = synth.unit_vectors(256, 128, 31)
    p = a + 0.05 * torch.randn(256, 128, generator=torch.Generator().manual_seed(32))
    p = p / p.norm(dim=1, keepdim=True)
    n = synth.unit_vectors(256, 128, 33)
    return a, p, n


def main():
    sys.path.insert(0, str(REF / "hardnet"))
    import Losses as ref  # the reference module itself
    torch.Tensor.cuda = lambda self, *a, **k: self   # loss_L2Net hard-codes .cuda() (Losses.py:65): device move only
    a, p, n = inputs()
    out = {}
    out["pair_ap"] = ref.distance_vectors_pairwise(a, p).numpy()
    d_ap, d_an, d_pn = ref.distance_vectors_pairwise(a, p, n)
    out["pair_an"], out["pair_pn"] = d_an.numpy(), d_pn.numpy()
    for lt in ("triplet_margin", "softmax", "contrastive"):
        for swap in (False, True):
            out[f"random_{lt}_swap{int(swap)}"] = np.array(ref.loss_random_sampling(a, p, n, anchor_swap=swap, margin=1.0, loss_type=lt).item())
    # loss_L2Net does `bool_tensor - 1` (Losses.py:70), which only old torch (ByteTensor masks) accepts. That mask is not
    # used by the softmax branch, the only branch the function implements; `ge` returning uint8 for the duration of the call
    # restores the old behaviour without touching the reference source.
    orig_ge = torch.Tensor.ge
    torch.Tensor.ge = lambda self, other: orig_ge(self, other).to(torch.uint8)
    try:
        for swap in (False, True):
            out[f"l2net_softmax_swap{int(swap)}"] = np.array(ref.loss_L2Net(a, p, anchor_swap=swap, loss_type="softmax").item())
    finally:
        torch.Tensor.ge = orig_ge
    out["gor"] = np.array(ref.global_orthogonal_regularization(a, n).item())
    np.savez_compressed(ROOT / "tests" / "golden" / "losses_extra.npz", **out)
    print({k: (float(v) if v.ndim == 0 else v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
