"""Per-stage time of one HardNet forward at small batches (CUDA events around every launch; BASELINE config 1 is 2 x 1024)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hardnetnas_b200.hardnet import HardNet
torch.manual_seed(0)
m = HardNet().cuda().eval()
for B in (256, 1024, 2048, 4096, 18944):
    x = torch.rand(B, 1, 32, 32, device="cuda")
    for _ in range(3):
        m(x)
    m.profile_enable(0x7F)
    for _ in range(10):
        m(x)
    torch.cuda.synchronize()
    ms, n = m.profile_read()
    m.profile_enable(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m(x)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B:6d} forward {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us | per stage us:",
          " ".join(f"{HardNet.STAGE_NAMES[i].split('_')[0]}={ms[i] / max(n[i], 1) * 1e3:.1f}" for i in range(1, 7)))
