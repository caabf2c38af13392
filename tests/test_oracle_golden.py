"""The CPU oracle against the golden vectors produced by the reference's own code (oracle/make_golden.py)."""
import numpy as np
import torch

from oracle import hardnet_oracle, losses_oracle, synth


def _load(golden_dir, name):
    return np.load(golden_dir / name)


def test_weights_regenerate_like_reference(golden_dir):
    g = _load(golden_dir, "hardnet_forward.npz")
    w, _, _ = synth.hardnet_weights_from_seed(0, 3)
    np.testing.assert_allclose(synth.weights_fingerprint(w), g["weights_fingerprint"], rtol=1e-5, atol=1e-6)
    x = synth.make_patches(64, 1234)
    chk = np.array([x.double().sum().item(), x[:, :, ::7, ::5].double().sum().item()])
    np.testing.assert_allclose(chk, g["patch_checksum"], rtol=1e-9)


def test_hardnet_forward_matches_reference(golden_dir):
    g = _load(golden_dir, "hardnet_forward.npz")
    w, m, v = synth.hardnet_weights_from_seed(0, 3)
    x = synth.make_patches(64, 1234)
    desc = hardnet_oracle.hardnet_forward(x, w, m, v).numpy()
    np.testing.assert_allclose(desc, g["desc"], atol=2e-6)
    np.testing.assert_allclose(hardnet_oracle.input_norm(x)[:2].numpy(), g["input_norm_first2"], atol=1e-5, rtol=1e-5)
    acts = hardnet_oracle.hardnet_stages(x, w, m, v)
    for n, a in enumerate(acts, start=1):
        mean = a.mean(dim=(2, 3)).numpy() if a.shape[-1] > 1 else a.reshape(a.size(0), -1).numpy()
        np.testing.assert_allclose(mean, g[f"stage{n}_mean"], atol=1e-4, rtol=1e-4)
        sample = a[:4].numpy() if n >= 5 else a[:2, :, ::4, ::4].numpy()
        np.testing.assert_allclose(sample, g[f"stage{n}_sample"], atol=2e-4, rtol=1e-4)


def test_hardnet_fresh_bn_constant_patch_is_zero(golden_dir):
    g = _load(golden_dir, "hardnet_forward.npz")
    w, m, v = synth.hardnet_weights_from_seed(0, None)
    x = synth.make_patches(64, 1234)
    desc = hardnet_oracle.hardnet_forward(x, w, m, v).numpy()
    np.testing.assert_allclose(desc, g["desc_fresh_bn"], atol=2e-6)
    assert np.all(g["desc_fresh_bn"][-1] == 0.0)  # constant patch -> exactly zero descriptor
    assert np.all(desc[-1] == 0.0)


def _loss_inputs():
    w, m, v = synth.hardnet_weights_from_seed(0, None)
    anchors = synth.make_patches(256, 1234)
    positives = synth.make_positives(anchors, 0.1, 7)
    da = hardnet_oracle.hardnet_forward(anchors, w, m, v)
    dp = hardnet_oracle.hardnet_forward(positives, w, m, v)
    return da, dp


def test_losses_match_reference(golden_dir):
    g = _load(golden_dir, "losses.npz")
    da, dp = _loss_inputs()
    np.testing.assert_allclose([da.double().sum().item(), da.double().abs().sum().item()], g["desc_a_checksum"], rtol=1e-5)
    dm = losses_oracle.distance_matrix_vector(da, dp)
    np.testing.assert_allclose(dm[:16, :16].numpy(), g["dist_matrix_16"], atol=2e-6)
    np.testing.assert_allclose(dm.double().sum(1).numpy(), g["dist_matrix_rowsum"], rtol=1e-5)
    for swap in (False, True):
        assert abs(losses_oracle.loss_hardnet(da, dp, swap, 1.0).item() - float(g[f"loss_swap{int(swap)}"])) < 2e-6
        assert abs(losses_oracle.loss_hardnet(da, dp, swap, 0.5).item() - float(g[f"loss_swap{int(swap)}_m05"])) < 2e-6
    dp2 = dp.clone()
    dp2[3] = da[3]
    dp2[5] = da[9]
    for swap in (False, True):
        assert abs(losses_oracle.loss_hardnet(da, dp2, swap).item() - float(g[f"loss_dup_swap{int(swap)}"])) < 2e-6


def test_losses_unit_vectors_match_reference(golden_dir):
    g = _load(golden_dir, "losses.npz")
    ua = synth.unit_vectors(512, 128, 21)
    up = synth.unit_vectors(512, 128, 22)
    up[:400] = ua[:400] + 0.05 * torch.randn(400, 128, generator=torch.Generator().manual_seed(23))
    up = up / up.norm(dim=1, keepdim=True)
    for swap in (False, True):
        assert abs(losses_oracle.loss_hardnet(ua, up, swap).item() - float(g[f"loss_unit_swap{int(swap)}"])) < 2e-6
    assert abs(losses_oracle.loss_hardnet(ua, up, True).item() - float(g["loss_nas_unit"])) < 2e-6


def test_fpr95_matches_reference(golden_dir):
    g = _load(golden_dir, "losses.npz")
    rng = np.random.RandomState(5)
    labels = (rng.rand(4000) < 0.5).astype(np.int64)
    dist = np.where(labels == 1, rng.gamma(2.0, 0.15, 4000), rng.gamma(6.0, 0.2, 4000)).astype(np.float32)
    assert losses_oracle.error_rate_at_95_recall(labels, 1.0 / (dist + 1e-8)) == float(g["fpr95"])


def test_matching_matches_reference(golden_dir):
    g = _load(golden_dir, "matching.npz")
    q, gal, truth = synth.make_match_set(768, 2048, seed=11)
    np.testing.assert_allclose([q.double().sum().item(), gal.double().sum().item()], g["q_checksum"], rtol=1e-9)
    np.testing.assert_array_equal(truth.numpy(), g["truth"])
    val, idx = losses_oracle.nn_match(q, gal, chunk=256)
    np.testing.assert_array_equal(idx.numpy(), g["nn_idx"])
    np.testing.assert_allclose(val.numpy(), g["nn_val"], atol=1e-6)
    lab, ia, da, db = losses_oracle.ratio_match(q, gal, 0.7, chunk=256)
    np.testing.assert_array_equal(lab.numpy(), g["ratio_label"])
    np.testing.assert_array_equal(ia.numpy(), g["ratio_idx"])
    np.testing.assert_allclose(db.numpy(), g["second_val"], atol=1e-6)
    np.testing.assert_allclose(losses_oracle.distance_matrix_vector_fdl(q, gal)[:16, :16].numpy(), g["dist_16"], atol=1e-6)
    # matched queries find their planted gallery row
    has = truth >= 0
    assert torch.equal(idx[has], truth[has])


def test_mutual_nn_definition():
    q, gal, truth = synth.make_match_set(300, 400, seed=5)
    pairs = losses_oracle.mutual_nn(q, gal)
    d = losses_oracle.distance_matrix_vector_fdl(q, gal)
    fwd, bwd = d.argmin(1), d.argmin(0)
    for i, j in pairs.tolist():
        assert fwd[i] == j and bwd[j] == i
    assert pairs.size(0) == int((bwd[fwd] == torch.arange(300)).sum())


def test_other_loss_helpers_match_reference_goldens(golden_dir):
    """distance_vectors_pairwise / loss_random_sampling / loss_L2Net / global_orthogonal_regularization (hardnet/Losses.py:
    15-85,156-162; imported by name at hardnet/HardNet.py:36) against values produced by the unmodified reference
    (oracle/make_golden_losses_extra.py). Host-side torch expressions, so they are checked on CPU tensors."""
    from oracle.make_golden_losses_extra import inputs
    from hardnetnas_b200 import losses
    g = _load(golden_dir, "losses_extra.npz")
    a, p, n = inputs()
    assert np.allclose(losses.distance_vectors_pairwise(a, p).numpy(), g["pair_ap"], atol=1e-6)
    d_ap, d_an, d_pn = losses.distance_vectors_pairwise(a, p, n)
    assert np.allclose(d_an.numpy(), g["pair_an"], atol=1e-6) and np.allclose(d_pn.numpy(), g["pair_pn"], atol=1e-6)
    for lt in ("triplet_margin", "softmax", "contrastive"):
        for swap in (False, True):
            got = losses.loss_random_sampling(a, p, n, anchor_swap=swap, margin=1.0, loss_type=lt).item()
            assert abs(got - float(g[f"random_{lt}_swap{int(swap)}"])) <= 1e-6, (lt, swap, got)
    for swap in (False, True):
        got = losses.loss_L2Net(a, p, anchor_swap=swap, loss_type="softmax").item()
        assert abs(got - float(g[f"l2net_softmax_swap{int(swap)}"])) <= 1e-5, (swap, got)
    assert abs(losses.global_orthogonal_regularization(a, n).item() - float(g["gor"])) <= 1e-9
