"""Small fixed NAS-net workload for ncu (run under gpurun)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hardnetnas_b200.nas import SampledDescriptorNet  # noqa: E402

torch.manual_seed(0)
net = SampledDescriptorNet(sys.argv[1] if len(sys.argv) > 1 else "wang2").cuda().eval()
x = torch.nn.functional.avg_pool2d(torch.rand(65536, 1, 32, 32, device="cuda"), 5, 1, 2)
for _ in range(2):
    net(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    net(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"done: {ms:.3f} ms per 65536 patches = {65536 / ms / 1e3:.2f} M patches/s")
