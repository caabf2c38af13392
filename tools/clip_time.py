"""Keypoints -> descriptors: hn_clip_patches + hn_forward (patch tensor through HBM) against hn_forward_clip (crop inside the front
kernel's loader warps), fp32 and uint8 images. CUDA events, 3 warm-up + 10 timed calls each.

    python tools/clip_time.py [n_keypoints_per_image] [images]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from hardnetnas_b200.hardnet import HardNet  # noqa: E402
from hardnetnas_b200.image_utils import clip_patch  # noqa: E402


def timed(fn, warm=3, rep=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep


k = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
H, W = 480, 640
g = torch.Generator().manual_seed(1)
images = torch.nn.functional.avg_pool2d(torch.rand(B, 1, H, W, generator=g), 5, 1, 2)
img8 = (images * 255).round().to(torch.uint8).cuda()
imgf = img8.float()
ys, xs = torch.randint(0, H // 2, (B, k), generator=g), torch.randint(0, W // 2, (B, k), generator=g)
bs = torch.arange(B)[:, None].expand(B, k)
byxc = torch.stack([bs, ys, xs, torch.zeros_like(ys)], dim=-1).view(-1, 4).long().cuda()
scale = (6.0 + 34.0 * torch.rand(B * k, generator=g)).cuda()
ang = 6.2831853 * torch.rand(B * k, generator=g)
ori = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).cuda()
info = torch.full((B, 2), 0.5).cuda()
torch.manual_seed(0)
model = HardNet().cuda().eval()
n = B * k
two = model(clip_patch(byxc, scale, ori, info, imgf, 32))
one = model.forward_clip(byxc, scale, ori, info, imgf)
one8 = model.forward_clip(byxc, scale, ori, info, img8)
print(f"{n} keypoints on {B} images of {H}x{W}: bit-identical fp32 {torch.equal(one, two)}, uint8 {torch.equal(one8, two)}")
patches = clip_patch(byxc, scale, ori, info, imgf, 32)
t_fwd = timed(lambda: model(patches))
t_two = timed(lambda: model(clip_patch(byxc, scale, ori, info, imgf, 32)))
t_one = timed(lambda: model.forward_clip(byxc, scale, ori, info, imgf))
t_one8 = timed(lambda: model.forward_clip(byxc, scale, ori, info, img8))
for name, t in (("forward of ready patches", t_fwd), ("clip_patch + forward", t_two), ("forward_clip fp32 images", t_one),
                ("forward_clip uint8 images", t_one8)):
    print(f"  {name:28s} {t:8.3f} ms  {n / t / 1e3:8.2f} M keypoints/s")
