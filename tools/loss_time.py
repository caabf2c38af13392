"""Where the time of the fused loss goes at small N (diagnostic): CUDA-graph replays of its pieces."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hardnetnas_b200 import _lib, _ops  # noqa: E402
from hardnetnas_b200.losses import loss_HardNet  # noqa: E402


def timeit(fn, n=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def graphed(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


lib = _lib.load()
for n in (128, 512, 1024, 2048, 4096, 8192):
    a = torch.nn.functional.normalize(torch.randn(n, 128, device="cuda"), dim=1)
    p = torch.nn.functional.normalize(a + 0.1 * torch.randn_like(a), dim=1)
    out16 = torch.empty(n * 384, dtype=torch.float16, device="cuda")
    stream = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    t_pack = timeit(graphed(lambda: lib.hn_pack_descriptors(a.data_ptr(), n, out16.data_ptr(), stream())))
    t_dist = timeit(graphed(lambda: _ops.dist_min(a, p, _lib.HN_FORM_HARDNET, True, True)))
    t_dist1 = timeit(graphed(lambda: _ops.dist_min(a, p, _lib.HN_FORM_HARDNET, True, False)))
    t_loss = timeit(graphed(lambda: loss_HardNet(a, p, anchor_swap=True)))
    t_empty = timeit(graphed(lambda: out16.zero_()))
    print(f"n={n}: one pack {t_pack:.1f} us, dist_min swap {t_dist:.1f} us, no swap {t_dist1:.1f} us, loss {t_loss:.1f} us, a memset node {t_empty:.1f} us",
          flush=True)
