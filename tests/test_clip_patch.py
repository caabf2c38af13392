"""clip_patch (FDLNet-master/utils/image_utils.py:11-158): oracle vs the reference-generated golden (CPU) and the
CUDA kernel behind hn_clip_patches vs the oracle (GPU)."""
import numpy as np
import pytest
import torch

from oracle import clip_oracle

TOL = 1e-5   # fp32 bilinear weights; the reference's stacked matmuls and the per-pixel form differ by ~2e-6


def test_oracle_matches_reference_golden(golden_dir):
    g = np.load(golden_dir / "clip_patch.npz")
    byxc, scale, ori, im_info, images = clip_oracle.make_clip_inputs()
    assert abs(images.double().sum().item() - float(g["images_sum"])) < 1e-6   # same seeded inputs
    out = clip_oracle.clip_patch(byxc, scale, ori, im_info, images, 32)
    assert out.shape == (byxc.size(0), 1, 32, 32)
    assert (out - torch.from_numpy(g["patches"])).abs().max().item() <= TOL
    out = clip_oracle.clip_patch(byxc, scale, None, im_info, images, 32)
    assert (out - torch.from_numpy(g["patches_no_ori"])).abs().max().item() <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("with_ori", [True, False])
def test_kernel_matches_oracle_and_golden(with_ori, golden_dir):
    from hardnetnas_b200.image_utils import clip_patch
    byxc, scale, ori, im_info, images = clip_oracle.make_clip_inputs()
    got = clip_patch(byxc.cuda(), scale.cuda(), ori.cuda() if with_ori else None, im_info.cuda(), images.cuda(), 32).cpu()
    ref = clip_oracle.clip_patch(byxc, scale, ori if with_ori else None, im_info, images, 32)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= TOL
    gold = torch.from_numpy(np.load(golden_dir / "clip_patch.npz")["patches" if with_ori else "patches_no_ori"])
    assert (got - gold).abs().max().item() <= TOL


@pytest.mark.gpu
def test_kernel_larger_set_and_feeds_descriptor():
    """2048 keypoints over 8 images (other seed, other patch size), then straight into HardNet.forward."""
    from hardnetnas_b200.hardnet import HardNet
    from hardnetnas_b200.image_utils import clip_patch
    byxc, scale, ori, im_info, images = clip_oracle.make_clip_inputs(seed=5, B=8, H=240, W=320, k=256)
    for psize in (32, 17):
        got = clip_patch(byxc.cuda(), scale.cuda(), ori.cuda(), im_info.cuda(), images.cuda(), psize).cpu()
        ref = clip_oracle.clip_patch(byxc, scale, ori, im_info, images, psize)
        assert (got - ref).abs().max().item() <= TOL
    torch.manual_seed(0)
    model = HardNet().cuda().eval()
    desc = model(clip_patch(byxc.cuda(), scale.cuda(), ori.cuda(), im_info.cuda(), images.cuda(), 32))
    assert desc.shape == (2048, 128) and torch.isfinite(desc).all()


@pytest.mark.gpu
def test_cpu_tensors_fail_loudly():
    from hardnetnas_b200._lib import HardnetB200Error
    from hardnetnas_b200.image_utils import clip_patch
    byxc, scale, ori, im_info, images = clip_oracle.make_clip_inputs()
    with pytest.raises(HardnetB200Error):
        clip_patch(byxc, scale, ori, im_info, images, 32)
