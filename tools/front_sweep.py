import os, subprocess, sys
for fc in (296, 592, 1184, 2368, 4736, 18944):
    env = dict(os.environ, HN_FRONT_CHUNK=str(fc))
    r = subprocess.run([sys.executable, "-c", """
import sys, time, torch
sys.path.insert(0, '.')
from hardnetnas_b200.hardnet import HardNet
torch.manual_seed(0)
m = HardNet().cuda().eval()
x = torch.nn.functional.avg_pool2d(torch.rand(262144, 1, 32, 32, device='cuda'), 5, 1, 2)
out = torch.empty(262144, 128, device='cuda')
for _ in range(3): m(x, out=out)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): m(x, out=out)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f'{262144/dt/1e6:.3f} M patches/s')
"""], env=env, capture_output=True, text=True)
    print("front_chunk", fc, r.stdout.strip(), r.stderr.strip()[-200:], flush=True)
