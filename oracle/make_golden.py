"""Generate tests/golden/*.npz by running the UNMODIFIED reference code from /root/reference.

Run in the build container only (the reference does not travel to the GPU box):

    python oracle/make_golden.py

Inputs come from oracle/synth.py (seeded), so the tests can regenerate them; the .npz files hold only the
reference's outputs plus small fingerprints of the regenerated inputs/weights.
"""
from __future__ import annotations

import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))

from oracle import synth  # noqa: E402

GOLDEN = REPO / "tests" / "golden"


def _import_reference_hardnet():
    """Import hardnet/HardNet.py safely: it parses argv, sets CUDA_VISIBLE_DEVICES and creates data/logs/
    in the cwd at import time (HardNet.py:56-134,154,166-167)."""
    saved_argv, saved_cwd = sys.argv, os.getcwd()
    saved_env = os.environ.get("CUDA_VISIBLE_DEVICES")
    tmp = tempfile.mkdtemp(prefix="hn_ref_")
    os.chdir(tmp)
    sys.argv = ["HardNet.py"]
    sys.path.insert(0, str(REF / "hardnet"))
    try:
        import HardNet as ref_hardnet  # noqa: N813
        import Losses as ref_losses
        import EvalMetrics as ref_metrics
    finally:
        sys.argv = saved_argv
        os.chdir(saved_cwd)
        sys.path.remove(str(REF / "hardnet"))
        if saved_env is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = saved_env
    return ref_hardnet, ref_losses, ref_metrics


def _import_reference_fdl():
    for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
        del sys.modules[m]
    sys.path.insert(0, str(REF / "FDLNet-master"))
    try:
        from utils import eval_utils, math_utils
    finally:
        sys.path.remove(str(REF / "FDLNet-master"))
    return eval_utils, math_utils


def golden_hardnet(ref_hardnet):
    torch.manual_seed(0)
    model = ref_hardnet.HardNet()
    sd = synth.randomize_bn_stats(model.state_dict(), seed=3)
    model.load_state_dict(sd)
    model.eval()

    # the oracle's weight generator must reproduce the reference constructor bit for bit
    w, means, vars_ = synth.hardnet_weights_from_seed(0, 3)
    for i, (ci, bi) in enumerate(zip(synth.CONV_IDX, synth.BN_IDX)):
        assert torch.equal(w[i], sd[f"features.{ci}.weight"]), f"weight {i} differs from the reference init"
        assert torch.equal(means[i], sd[f"features.{bi}.running_mean"])
        assert torch.equal(vars_[i], sd[f"features.{bi}.running_var"])

    x = synth.make_patches(64, seed=1234)
    with torch.no_grad():
        desc = model(x)
        normed = model.input_norm(x)
        # per-stage activations through the reference's own nn.Sequential
        h = normed
        stage_out = {}
        for idx, layer in enumerate(model.features):
            h = layer(h)
            if idx in (2, 5, 8, 11, 14, 17, 20):
                stage_out[idx] = h.clone()
    # default (fresh) BN stats model for the zero-descriptor edge case
    torch.manual_seed(0)
    model0 = ref_hardnet.HardNet().eval()
    with torch.no_grad():
        desc_fresh = model0(x)

    stage_keys = [2, 5, 8, 11, 14, 17, 20]
    out = {
        "desc": desc.numpy(),
        "desc_fresh_bn": desc_fresh.numpy(),
        "input_norm_first2": normed[:2].numpy(),
        "weights_fingerprint": synth.weights_fingerprint(w),
        "patch_checksum": np.array([x.double().sum().item(), x[:, :, ::7, ::5].double().sum().item()]),
    }
    for n, k in enumerate(stage_keys, start=1):
        a = stage_out[k]
        # keep the goldens small: per-patch channel means and a strided sample of each stage
        out[f"stage{n}_mean"] = a.mean(dim=(2, 3)).numpy() if a.dim() == 4 and a.shape[-1] > 1 else a.reshape(a.size(0), -1).numpy()
        out[f"stage{n}_sample"] = a[:4].numpy() if n >= 5 else a[:2, :, ::4, ::4].numpy()
    np.savez_compressed(GOLDEN / "hardnet_forward.npz", **out)
    print("hardnet_forward.npz", {k: v.shape for k, v in out.items()})
    return model, model0


def golden_losses(ref_losses, ref_metrics, model):
    torch.Tensor.cuda = lambda self, *a, **k: self  # Losses.py:96 hard-codes .cuda(); device move only
    anchors = synth.make_patches(256, seed=1234)
    positives = synth.make_positives(anchors, 0.1, 7)
    with torch.no_grad():
        da, dp = model(anchors), model(positives)
    out = {}
    out["desc_a_checksum"] = np.array([da.double().sum().item(), da.double().abs().sum().item()])
    dm = ref_losses.distance_matrix_vector(da, dp)
    out["dist_matrix_16"] = dm[:16, :16].numpy()
    out["dist_matrix_rowsum"] = dm.double().sum(1).numpy()
    for swap in (False, True):
        out[f"loss_swap{int(swap)}"] = np.array(ref_losses.loss_HardNet(da, dp, anchor_swap=swap, margin=1.0).item())
        out[f"loss_swap{int(swap)}_m05"] = np.array(ref_losses.loss_HardNet(da, dp, anchor_swap=swap, margin=0.5).item())
    # duplicated rows exercise the (<0.008 -> +10) mask: positives 3 and 5 are copies of anchors 3 / 9
    dp2 = dp.clone()
    dp2[3] = da[3]
    dp2[5] = da[9]
    for swap in (False, True):
        out[f"loss_dup_swap{int(swap)}"] = np.array(ref_losses.loss_HardNet(da, dp2, anchor_swap=swap).item())
    # unit-vector inputs (independent of the conv stack)
    ua = synth.unit_vectors(512, 128, 21)
    up = synth.unit_vectors(512, 128, 22)
    up[:400] = ua[:400] + 0.05 * torch.randn(400, 128, generator=torch.Generator().manual_seed(23))
    up = up / up.norm(dim=1, keepdim=True)
    for swap in (False, True):
        out[f"loss_unit_swap{int(swap)}"] = np.array(ref_losses.loss_HardNet(ua, up, anchor_swap=swap).item())
    # NAS variant (always swap), hardnetNAS/general_functions/Losses.py
    sys.path.insert(0, str(REF / "hardnetNAS"))
    for m in [k for k in sys.modules if k.startswith("general_functions")]:
        del sys.modules[m]
    import importlib.util
    spec = importlib.util.spec_from_file_location("nas_losses", REF / "hardnetNAS/general_functions/Losses.py")
    nas_losses = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nas_losses)
    sys.path.remove(str(REF / "hardnetNAS"))
    out["loss_nas_unit"] = np.array(nas_losses.loss_HardNet(ua, up, margin=1.0).item())

    # FPR95 (EvalMetrics.py) on a seeded pair set
    rng = np.random.RandomState(5)
    labels = (rng.rand(4000) < 0.5).astype(np.int64)
    dist = np.where(labels == 1, rng.gamma(2.0, 0.15, 4000), rng.gamma(6.0, 0.2, 4000)).astype(np.float32)
    out["fpr95"] = np.array(ref_metrics.ErrorRateAt95Recall(labels, 1.0 / (dist + 1e-8)))
    np.savez_compressed(GOLDEN / "losses.npz", **out)
    print("losses.npz", {k: (v.shape, float(v) if v.ndim == 0 else None) for k, v in out.items()})


def golden_matching(eval_utils, math_utils):
    q, g, truth = synth.make_match_set(768, 2048, seed=11)
    d = math_utils.distance_matrix_vector(q, g)
    nn_val, nn_idx = d.min(dim=-1)
    kp2 = torch.arange(g.size(0)).view(-1, 1).float()
    label, nn_kp2 = eval_utils.nearest_neighbor_distance_ratio_match(q, g, kp2, 0.7)
    srt, _ = d.sort(dim=-1)
    out = {
        "nn_val": nn_val.numpy(),
        "nn_idx": nn_idx.numpy(),
        "ratio_label": label.numpy(),
        "ratio_idx": nn_kp2.view(-1).long().numpy(),
        "second_val": srt[:, 1].numpy(),
        "dist_16": d[:16, :16].numpy(),
        "truth": truth.numpy(),
        "q_checksum": np.array([q.double().sum().item(), g.double().sum().item()]),
    }
    np.savez_compressed(GOLDEN / "matching.npz", **out)
    print("matching.npz", {k: v.shape for k, v in out.items()})


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref_hardnet, ref_losses, ref_metrics = _import_reference_hardnet()
    _, model_fresh = golden_hardnet(ref_hardnet)
    # fresh BatchNorm statistics give well-spread descriptors (d ~ 0.5); with randomised statistics they are
    # nearly parallel (d ~ 0.02) and the reference formula itself is noise-dominated
    golden_losses(ref_losses, ref_metrics, model_fresh)
    eval_utils, math_utils = _import_reference_fdl()
    golden_matching(eval_utils, math_utils)
    if "--nas" in sys.argv or True:
        try:
            from oracle import make_golden_nas
            make_golden_nas.main()
        except ImportError:
            print("NAS golden generator not present yet")


if __name__ == "__main__":
    main()
