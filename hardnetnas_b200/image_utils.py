"""Drop-in mirror of FDLNet-master/utils/image_utils.py:clip_patch — the step right upstream of the descriptor
(SURVEY.md section 8f row 2): keypoint (b, y, x) + scale + orientation -> bilinear PSIZE x PSIZE crop.

Same signature and return shape as the reference; runs as one CUDA kernel behind the C ABI (hn_clip_patches).
CUDA tensors only; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def clip_patch(kpts_byxc, kpts_scale, kpts_ori, im_info, images, PSIZE):
    """clip patch from the raw images (image_utils.py:11-158).

    kpts_byxc [N,4] integer (b, y, x, 0); kpts_scale [N]; kpts_ori [N,2] (cos, sin) or None; im_info [B,2];
    images [B,1,H,W]; returns [N,1,PSIZE,PSIZE] fp32.
    """
    assert kpts_byxc.size(0) == kpts_scale.size(0)   # image_utils.py:22
    if not (isinstance(images, torch.Tensor) and images.is_cuda):
        raise _lib.HardnetB200Error("clip_patch runs on B200 CUDA tensors only (no CPU fallback)")
    dev = images.device
    B, Cc, H, W = images.size()
    if Cc != 1:
        raise ValueError("clip_patch expects single-channel images [B,1,H,W] (the reference flattens them as such)")
    n = kpts_byxc.size(0)
    img = images.detach().to(torch.float32).contiguous()
    byxc = kpts_byxc.detach().to(device=dev, dtype=torch.int64).contiguous()
    scale = kpts_scale.detach().to(device=dev, dtype=torch.float32).contiguous().view(-1)
    ori = None if kpts_ori is None else kpts_ori.detach().to(device=dev, dtype=torch.float32).contiguous()
    info = im_info.detach().to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty((n, 1, PSIZE, PSIZE), dtype=torch.float32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.hn_clip_patches(img.data_ptr(), B, H, W, byxc.data_ptr(), scale.data_ptr(),
                                       None if ori is None else ori.data_ptr(), info.data_ptr(), n, int(PSIZE),
                                       out.data_ptr(), C.c_void_p(stream)), "hn_clip_patches")
    return out
