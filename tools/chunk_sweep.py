"""Throughput of HardNet.forward vs conv-stack pass size (diagnostic; run on the GPU box)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hardnetnas_b200.hardnet import HardNet  # noqa: E402

N = 262144
x = torch.nn.functional.avg_pool2d(torch.rand(N, 1, 32, 32, device="cuda"), 5, 1, 2)
out = torch.empty(N, 128, device="cuda")
for chunk in [int(a) for a in sys.argv[1:]] or [1184, 2368, 4736, 9472, 18944]:
    torch.manual_seed(0)
    m = HardNet(chunk_patches=chunk, head_rows=max(chunk, 18944)).cuda().eval()
    for _ in range(3):
        m(x, out=out)
    torch.cuda.synchronize()
    m.profile_enable(0x7F)
    m(x, out=out)
    torch.cuda.synchronize()
    ms, _ = m.profile_read()
    m.profile_enable(0)
    t0 = time.perf_counter()
    for _ in range(5):
        m(x, out=out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk {chunk}: {N / dt / 1e6:.3f} M patches/s; per-stage ns/patch {[round(v * 1e6 / N, 1) for v in ms]}", flush=True)
    del m
