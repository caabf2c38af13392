"""A/B timing of the matching calls at BASELINE config 4 size (64k x 64k, oracle/synth.make_match_set(seed=11))."""
import sys
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from hardnetnas_b200 import _ops  # noqa: E402
from hardnetnas_b200.matching import mutual_nn_ratio, mutual_nn_ratio_two_pass  # noqa: E402
from oracle import synth  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
q, g, _ = synth.make_match_set(n, n, seed=11)
q, g = q.cuda(), g.cuda()
q16, g16 = _ops.pack_descriptors(q), _ops.pack_descriptors(g)
bm = torch.empty(_ops.block_max_elems(n, n), dtype=torch.float32, device="cuda")
print(f"n={n}")
print("forward only (packs inside)      %.3f ms" % timeit(lambda: _ops.match_top2(q, g)))
print("forward only, prepacked          %.3f ms" % timeit(lambda: _ops.match_top2(q, g, q16=q16, g16=g16)))
print("forward + block maxima, prepacked %.3f ms" % timeit(lambda: _ops.match_top2(q, g, q16=q16, g16=g16, block_max=bm)))
print("mutual, single GEMM              %.3f ms" % timeit(lambda: mutual_nn_ratio(q, g, return_pairs=False)))
print("mutual, two GEMM passes          %.3f ms" % timeit(lambda: mutual_nn_ratio_two_pass(q, g, return_pairs=False)))
a = mutual_nn_ratio(q, g, return_pairs=False)
b = mutual_nn_ratio_two_pass(q, g, return_pairs=False)
print("equal:", torch.equal(a[1], b[1]), torch.equal(a[3], b[3]), "mutual rows:", int(a[1].sum()))
