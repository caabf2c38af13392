"""Diagnostic: where each role of the fused front kernel waits (needs a build with HN_EXTRA_NVCC_FLAGS=-DHN_FF_TRACE)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hardnetnas_b200 import _lib  # noqa: E402
from hardnetnas_b200.hardnet import HardNet  # noqa: E402

import os  # noqa: E402
if os.environ.get("HN_TRACE_LIB"):   # a trace build next to the normal library (tools/c34_trace.py --build makes one with both trace flags)
    _lib.LIB_PATH = Path(os.environ["HN_TRACE_LIB"])
lib = _lib.load()
lib.hn_debug_ff_trace.restype = C.c_int
lib.hn_debug_ff_trace.argtypes = [C.c_void_p, C.c_int]
torch.manual_seed(0)
x = torch.nn.functional.avg_pool2d(torch.rand(18944 * 4, 1, 32, 32, device="cuda"), 5, 1, 2)
if len(sys.argv) > 1 and sys.argv[1] == "nas":       # the PW2 variant (NAS stem + first pointwise conv)
    from hardnetnas_b200.nas import SampledDescriptorNet
    net = SampledDescriptorNet("wang2").cuda().eval()
    run = lambda: net(x)
else:
    m = HardNet().cuda().eval()
    out = torch.empty(x.size(0), 128, device="cuda")
    run = lambda: m(x, out=out)
for _ in range(2):
    run()
buf = (C.c_ulonglong * 32)()
lib.hn_debug_ff_trace(buf, 1)
run()
lib.hn_debug_ff_trace(buf, 0)
v = list(buf)
patches = max(v[16], 1)
names = {0: "loader: raw_full (x4 warps)", 1: "loader: a1_empty (x4)", 2: "issuer: a1_full", 3: "issuer: l1_empty (2nd half)",
         4: "issuer: act1_full", 5: "issuer: l1_empty (1st half)", 6: "issuer: c2_empty (8 tiles)", 7: "issuer: mma_done (8 tiles)",
         8: "L1 epi: act1_empty (x4 warps)", 9: "L1 epi: l1_full (x4, 2 halves)", 10: "conv2 epi: c2_full (x16 warps... per group tiles)",
         11: "issuer: stage-1 MMA issue sections", 12: "issuer: PW2 conv2 MMA issue sections",
         13: "issuer: conv2 MMA issue sections (8 tiles)", 14: "FDW: depthwise phase (x8 warps)",
         17: "FDW: barrier before the depthwise phase (x8)", 18: "FDW: barrier behind it (x8)"}
print(f"CTA 0: {v[15] / patches:.0f} cycles per patch over {patches} patches")
for k, n in names.items():
    print(f"  {n:50s} {v[k] / patches:9.0f} cycles/patch")
