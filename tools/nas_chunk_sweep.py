import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hardnetnas_b200.nas import SampledDescriptorNet
from hardnetnas_b200.hardnet import HardNet

x = torch.nn.functional.avg_pool2d(torch.rand(65536, 1, 32, 32, device="cuda"), 5, 1, 2)
def bench(net, tag):
    for _ in range(2): net(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): net(x)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"{tag}: {65536/dt/1e6:.3f} M patches/s", flush=True)
for chunk in (9472, 18944, 37888):
    torch.manual_seed(0)
    bench(SampledDescriptorNet("wang2", chunk_patches=chunk, head_rows=37888).cuda().eval(), f"wang2 chunk {chunk}")
for chunk in (2368, 3552, 4736, 9472, 18944):
    torch.manual_seed(0)
    bench(HardNet(chunk_patches=chunk, head_rows=37888).cuda().eval(), f"hardnet chunk {chunk}")
