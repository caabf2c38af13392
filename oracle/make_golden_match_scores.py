"""Golden values for the match-score counters of FDLNet-master/utils/eval_utils.py (nearest_neighbor_match_score :112-127,
nearest_neighbor_threshold_match_score :130-150, nearest_neighbor_distance_ratio_match_score :178-197), produced by
importing the UNMODIFIED reference in this container. TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_match_scores.py      (needs /root/reference; writes tests/golden/match_scores.npz)
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import synth  # noqa: E402

REF = Path("/root/reference")
DES_THRSH, COO_THRSH = 1.0, 5.0


def inputs():
    """Descriptors of oracle.synth.make_match_set plus keypoints: kp2 = random (b, y, x) per gallery row, kp1w = the true
    match's keypoint + small noise (random position for distractor queries), visible = random 85 % mask."""
    q, g, truth = synth.make_match_set(768, 2048, seed=11)
    gen = torch.Generator().manual_seed(41)
    kp2 = torch.cat([torch.zeros(g.size(0), 1), torch.rand(g.size(0), 2, generator=gen) * 480.0], dim=1)
    kp1w = torch.cat([torch.zeros(q.size(0), 1), torch.rand(q.size(0), 2, generator=gen) * 480.0], dim=1)
    has = truth >= 0
    kp1w[has, 1:3] = kp2[truth[has], 1:3] + 3.0 * torch.randn(int(has.sum()), 2, generator=gen)
    visible = torch.rand(q.size(0), generator=gen) < 0.85
    return q, g, kp1w, kp2, visible


def main():
    sys.path.insert(0, str(REF / "FDLNet-master"))
    from utils import eval_utils, math_utils
    q, g, kp1w, kp2, visible = inputs()
    out = {
        "nn": np.array(eval_utils.nearest_neighbor_match_score(q, g, kp1w, kp2, visible, COO_THRSH)),
        "nn_thresh": np.array(eval_utils.nearest_neighbor_threshold_match_score(q, g, kp1w, kp2, visible, DES_THRSH, COO_THRSH)),
        "nn_ratio": np.array(eval_utils.nearest_neighbor_distance_ratio_match_score(q, g, kp1w, kp2, visible, COO_THRSH)),
    }
    out["pairwise_16"] = math_utils.pairwise_distances(kp1w[:16, 1:3], kp2[:24, 1:3]).numpy()
    out["pairwise_self_16"] = math_utils.pairwise_distances(kp1w[:16, 1:3]).numpy()
    np.savez_compressed(ROOT / "tests" / "golden" / "match_scores.npz", **out)
    print({k: (v.tolist() if v.size < 8 else v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
