"""GPU parity of fused brute-force matching (NN, ratio test, mutual NN) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import losses_oracle, synth

pytestmark = pytest.mark.gpu

TIE_TAU = 1e-5
DIST_TOL = 2e-6


@pytest.fixture
def force_kernel():
    """hn_match_force_kernel(mode): -1 by size, 0 single-CTA GEMM, 1 CTA-pair GEMM; restored afterwards."""
    from hardnetnas_b200 import _lib
    lib = _lib.load()
    yield lambda mode: _lib.check(lib.hn_match_force_kernel(int(mode)), "hn_match_force_kernel")
    lib.hn_match_force_kernel(-1)


def _check(q, g, chunk=1024):
    from hardnetnas_b200.matching import match_top2
    d1, d2, i1, i2 = [t.cpu() for t in match_top2(q.cuda(), g.cuda())]
    lab, ia, da, db = losses_oracle.ratio_match(q, g, 0.7, chunk) if g.size(0) >= 2 else (None, None, None, None)
    if g.size(0) >= 2:
        assert (d1 - da).abs().max().item() <= DIST_TOL
        assert (d2 - db).abs().max().item() <= DIST_TOL
        bad = (i1 != ia).nonzero().flatten().tolist()
        if bad:
            d = losses_oracle.distance_matrix_vector_fdl(q[bad], g)
            for r, i in enumerate(bad):
                assert abs(d[r, i1[i]].item() - d[r, ia[i]].item()) <= TIE_TAU, f"row {i}: {i1[i]} vs {ia[i]}"
        got_lab = (d1 / d2) < 0.7
        diff = (got_lab != lab).nonzero().flatten().tolist()
        for i in diff:
            assert abs((da[i] / db[i]).item() - 0.7) <= 1e-4
    else:
        v, i = losses_oracle.nn_match(q, g, chunk)
        assert (d1 - v).abs().max().item() <= DIST_TOL and torch.equal(i1, i)
    return d1, d2, i1, i2


def test_matches_reference_goldens(golden_dir):
    from hardnetnas_b200.matching import nearest_neighbor_distance_ratio_match, nearest_neighbor_match
    g = np.load(golden_dir / "matching.npz")
    q, gal, truth = synth.make_match_set(768, 2048, seed=11)
    val, idx = nearest_neighbor_match(q.cuda(), gal.cuda())
    np.testing.assert_array_equal(idx.cpu().numpy(), g["nn_idx"])
    np.testing.assert_allclose(val.cpu().numpy(), g["nn_val"], atol=DIST_TOL)
    kp2 = torch.arange(2048).view(-1, 1).float().cuda()
    lab, nn_kp2 = nearest_neighbor_distance_ratio_match(q.cuda(), gal.cuda(), kp2, 0.7)
    np.testing.assert_array_equal(lab.cpu().numpy(), g["ratio_label"])
    np.testing.assert_array_equal(nn_kp2.view(-1).long().cpu().numpy(), g["ratio_idx"])


@pytest.mark.parametrize("nq,ng", [(1, 1), (1, 2), (5, 7), (8, 8), (100, 9), (256, 128), (257, 129), (1000, 4097), (4096, 8192)])
def test_ragged_sizes(nq, ng):
    q, g, _ = synth.make_match_set(nq, ng, seed=3 + nq)
    _check(q, g)


@pytest.mark.parametrize("nq,ng", [(1, 1), (5, 7), (257, 129), (511, 64), (513, 4097), (1000, 130), (4096, 8192), (9000, 20000)])
def test_ragged_sizes_cta_pair_kernel(nq, ng, force_kernel):
    """Same checks with the CTA-pair (cta_group::2) matching kernel forced on (by default it is used from 8192 queries)."""
    force_kernel(1)
    q, g, _ = synth.make_match_set(nq, ng, seed=3 + nq)
    _check(q, g)


def test_single_and_pair_kernels_agree(force_kernel):
    from hardnetnas_b200.matching import match_top2
    q, g, _ = synth.make_match_set(3000, 9000, seed=8)
    outs = []
    for mode in (0, 1):
        force_kernel(mode)
        outs.append([t.cpu() for t in match_top2(q.cuda(), g.cuda())])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_unstructured_descriptors_near_ties():
    # worst case for a 16-bit shortlist: all distances close together
    q = synth.unit_vectors(2048, 128, 5)
    g = synth.unit_vectors(6000, 128, 6)
    _check(q, g)


def test_duplicate_gallery_rows_pick_first_index():
    q, g, _ = synth.make_match_set(64, 512, seed=9)
    g[300] = g[20]
    g[40] = g[20]
    q[0] = g[20]
    d1, d2, i1, i2 = _check(q, g)
    assert i1[0].item() == 20 and i2[0].item() == 40


@pytest.mark.parametrize("nq,ng,pair", [(3000, 5000, -1), (5000, 3000, 1), (777, 1234, 0), (9000, 9000, -1), (33, 8, -1), (4100, 2050, 1)])
def test_mutual_nn_single_gemm_matches_oracle_and_two_pass(nq, ng, pair, force_kernel):
    """hn_match_mutual (one GEMM: block maxima -> claims -> exact verification) against the CPU oracle and against the
    two-pass composition, on both matching kernels and ragged sizes."""
    from hardnetnas_b200.matching import mutual_nn_ratio, mutual_nn_ratio_two_pass
    force_kernel(pair)
    q, g, _ = synth.make_match_set(nq, ng, seed=21)
    pairs, mutual, ratio, fwd, d1, d2 = mutual_nn_ratio(q.cuda(), g.cuda())
    ref = losses_oracle.mutual_nn(q, g)
    assert torch.equal(pairs.cpu(), ref), (nq, ng)
    p2, m2, r2, f2, a2, b2 = mutual_nn_ratio_two_pass(q.cuda(), g.cuda())
    assert torch.equal(mutual, m2) and torch.equal(ratio, r2) and torch.equal(fwd, f2) and torch.equal(d1, a2)


def test_mutual_nn_duplicate_rows_and_unstructured_sets():
    """Tie rule (lowest index wins in both directions) and a set without planted structure (many weak claims)."""
    from hardnetnas_b200.matching import mutual_nn_ratio, mutual_nn_ratio_two_pass
    g = synth.unit_vectors(4096, 128, 5)
    q = synth.unit_vectors(2048, 128, 6)
    q[10] = g[100]; q[11] = g[100]          # two identical queries claim the same column: the lower row is mutual
    g[200] = g[300]                          # duplicate gallery rows: the query picks the lower column
    q[12] = g[300]
    pairs, mutual, _, fwd, _, _ = mutual_nn_ratio(q.cuda(), g.cuda())
    assert mutual[10].item() and not mutual[11].item() and fwd[10].item() == 100 and fwd[11].item() == 100
    assert fwd[12].item() == 200 and mutual[12].item()
    assert torch.equal(pairs.cpu(), losses_oracle.mutual_nn(q, g))
    assert torch.equal(mutual, mutual_nn_ratio_two_pass(q.cuda(), g.cuda())[1])


def test_mutual_nn_matches_oracle():
    from hardnetnas_b200.matching import mutual_nearest_neighbors
    q, g, truth = synth.make_match_set(3000, 5000, seed=21)
    pairs = mutual_nearest_neighbors(q.cuda(), g.cuda()).cpu()
    ref = losses_oracle.mutual_nn(q, g)
    assert torch.equal(pairs, ref)
    planted = truth >= 0
    got = torch.full((3000,), -1, dtype=torch.long)
    got[pairs[:, 0]] = pairs[:, 1]
    assert (got[planted] == truth[planted]).float().mean().item() > 0.99


def test_config4_full_size_properties():
    """BASELINE config 4 (64k x 64k): planted matches are recovered, a row sample equals the oracle, and the
    mutual check is self-consistent."""
    from hardnetnas_b200.matching import match_top2
    q, g, truth = synth.make_match_set(65536, 65536, seed=11)
    qc, gc = q.cuda(), g.cuda()
    d1, d2, i1, i2 = match_top2(qc, gc)
    planted = (truth >= 0)
    assert torch.equal(i1.cpu()[planted], truth[planted])
    assert (d1 <= d2).all()
    rows = torch.arange(0, 65536, 32)
    lab, ia, da, db = losses_oracle.ratio_match(q[rows], g, 0.7, chunk=512)
    assert (d1.cpu()[rows] - da).abs().max().item() <= DIST_TOL
    assert (d2.cpu()[rows] - db).abs().max().item() <= DIST_TOL
    mism = (i1.cpu()[rows] != ia).nonzero().flatten().tolist()
    for r in mism:
        assert abs(d1.cpu()[rows][r].item() - da[r].item()) <= TIE_TAU
    # ratio test rejects the distractor queries
    ratio = (d1 / d2).cpu()
    assert (ratio[planted] < 0.7).float().mean().item() > 0.99
    assert (ratio[~planted] < 0.7).float().mean().item() < 0.01
    # mutual NN at full size from the single GEMM: equal to the two-pass composition, planted pairs are mutual
    from hardnetnas_b200.matching import mutual_nn_ratio, mutual_nn_ratio_two_pass
    _, mutual, _, fwd, _, _ = mutual_nn_ratio(qc, gc, return_pairs=False)
    _, m2, _, f2, _, _ = mutual_nn_ratio_two_pass(qc, gc, return_pairs=False)
    assert torch.equal(fwd, f2) and torch.equal(mutual, m2)
    assert mutual.cpu()[planted].float().mean().item() > 0.999


def test_match_score_counters_match_reference_goldens(golden_dir):
    """nearest_neighbor_match_score / _threshold_match_score / _distance_ratio_match_score (FDLNet-master/utils/eval_utils.py:
    112-197) on the fused matching kernel against the counters the unmodified reference produced
    (oracle/make_golden_match_scores.py)."""
    import numpy as np
    from hardnetnas_b200 import matching
    from oracle.make_golden_match_scores import COO_THRSH, DES_THRSH, inputs
    g = np.load(golden_dir / "match_scores.npz")
    q, gal, kp1w, kp2, visible = (t.cuda() for t in inputs())
    assert list(matching.nearest_neighbor_match_score(q, gal, kp1w, kp2, visible, COO_THRSH)) == g["nn"].tolist()
    assert list(matching.nearest_neighbor_threshold_match_score(q, gal, kp1w, kp2, visible, DES_THRSH, COO_THRSH)) == g["nn_thresh"].tolist()
    assert list(matching.nearest_neighbor_distance_ratio_match_score(q, gal, kp1w, kp2, visible, COO_THRSH)) == g["nn_ratio"].tolist()
