"""Per-op cycle counts of the NAS tail kernel (csrc/nas_tail.cuh) as seen by warpgroup 0 of CTA 0. Diagnostic build:

    HN_EXTRA_NVCC_FLAGS=-DHN_TAIL_TRACE python -m hardnetnas_b200.build --force && python tools/nas_tail_trace.py wang2
"""
import ctypes as C
import os
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from hardnetnas_b200 import _lib  # noqa: E402
from hardnetnas_b200.nas import SampledDescriptorNet  # noqa: E402

KN = {0: "STEM", 1: "PW", 2: "DW", 3: "MAXPOOL", 4: "SE", 5: "HEAD"}


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "wang2"
    lib = _lib.load()
    fn = lib.hn_debug_tail_trace
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    torch.manual_seed(0)
    net = SampledDescriptorNet(arch).cuda().eval()
    x = torch.rand((18944, 1, 32, 32), device="cuda")
    net(x)
    buf = (C.c_ulonglong * 128)()
    fn(buf, 1)
    net(x)
    fn(buf, 1)
    t = list(buf)
    prog = net.compile_program()
    plan = net.resident_plan()
    print("env", {k: v for k, v in os.environ.items() if k.startswith("HN_")}, "plan", plan)
    for li, (a, b, nwg, _) in enumerate(plan):
        n = max(t[56 + li], 1)
        print(f"launch {li}: ops {a}..{b} warpgroups={nwg} patches(WG0 of CTA0)={n} total/patch={t[48 + li] / n:.0f} clk "
              f"load={t[32 + li] / n:.0f} store={t[40 + li] / n:.0f}  => {t[48 + li] / n / nwg:.0f} clk/patch/SM")
        for i in range(a, b + 1):
            o = prog.ops[i]
            extra = f" (accumulators complete at {t[64 + i] / n:.0f})" if o.kind == 1 else ""
            print(f"   op {i:2d} {KN[o.kind]:8s} {o.cin:3d}->{o.cout:3d} k{o.kernel} s{o.stride} {o.hin:2d}->{o.hout:2d}: {t[i] / n:.0f} clk{extra}")


if __name__ == "__main__":
    main()
