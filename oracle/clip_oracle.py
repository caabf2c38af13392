"""CPU restatement of FDLNet-master/utils/image_utils.py:clip_patch (test infrastructure only: imported by tests/,
never by the product package). Pinned by tests/golden/clip_patch.npz, which oracle/make_golden_clip.py generates by
executing the reference's own function source."""
from __future__ import annotations

import torch


def clip_patch(kpts_byxc, kpts_scale, kpts_ori, im_info, images, psize: int):
    """image_utils.py:11-158, per output pixel instead of through stacked matmuls / gathers."""
    B, _, H, W = images.shape
    n = kpts_byxc.size(0)
    lin = torch.linspace(-1, 1, psize, dtype=torch.float32)                      # :30-35
    yt, xt = torch.meshgrid([lin, lin], indexing="ij")
    ratio = im_info[:, 0].float().repeat_interleave(n // B)                      # view(B, -1) / im_info, :53-55, 82-88
    s = (kpts_scale.float().view(-1) / ratio) / 2.0                              # :55-56
    if kpts_ori is not None:                                                     # thetas @ R, :67-73
        c, sn = kpts_ori[:, 0].float(), kpts_ori[:, 1].float()
        t00, t01, t10, t11 = s * c, s * (-sn), s * sn, s * c
    else:
        t00, t01, t10, t11 = s, torch.zeros_like(s), torch.zeros_like(s), s
    x = t00[:, None, None] * xt + t01[:, None, None] * yt                        # :77-79
    y = t10[:, None, None] * xt + t11[:, None, None] * yt
    x = x + (kpts_byxc[:, 2].float() / ratio)[:, None, None]                     # :82-93
    y = y + (kpts_byxc[:, 1].float() / ratio)[:, None, None]
    x0u, y0u = x.floor().long(), y.floor().long()                                # :98-106
    x0, x1 = x0u.clamp(0, W - 1), (x0u + 1).clamp(0, W - 1)
    y0, y1 = y0u.clamp(0, H - 1), (y0u + 1).clamp(0, H - 1)
    img = images[:, 0].float()
    b = kpts_byxc[:, 0].long()[:, None, None].expand_as(x0)
    Ia, Ib, Ic, Id = img[b, y0, x0], img[b, y1, x0], img[b, y0, x1], img[b, y1, x1]   # :126-138
    x0f, x1f, y0f, y1f = x0.float(), x1.float(), y0.float(), y1.float()
    wa, wb = (x1f - x) * (y1f - y), (x1f - x) * (y - y0f)                        # :146-149
    wc, wd = (x - x0f) * (y1f - y), (x - x0f) * (y - y0f)
    return (wa * Ia + wb * Ib + wc * Ic + wd * Id).unsqueeze(1)                  # :151-158


def make_clip_inputs(seed: int = 21, B: int = 3, H: int = 120, W: int = 160, k: int = 16):
    """Seeded synthetic inputs: smooth images, keypoints incl. image borders, scales 6..40 px, unit orientations."""
    g = torch.Generator().manual_seed(seed)
    images = torch.nn.functional.avg_pool2d(torch.rand(B, 1, H, W, generator=g), 5, 1, 2)
    ys = torch.randint(0, H // 2, (B, k), generator=g)
    xs = torch.randint(0, W // 2, (B, k), generator=g)
    ys[:, 0], xs[:, 0] = 0, 0                       # corner / border keypoints exercise the clamped taps
    ys[:, 1], xs[:, 1] = H // 2 - 1, W // 2 - 1
    bs = torch.arange(B)[:, None].expand(B, k)
    byxc = torch.stack([bs, ys, xs, torch.zeros_like(ys)], dim=-1).view(-1, 4).long()
    scale = (6.0 + 34.0 * torch.rand(B * k, generator=g))
    ang = 6.2831853 * torch.rand(B * k, generator=g)
    ori = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1)
    im_info = torch.full((B, 2), 0.5)               # rescale ratio sh, sw (image_utils.py:17)
    return byxc, scale, ori, im_info, images
