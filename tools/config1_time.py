"""BASELINE config 1 step (forward x 2 + loss at batch 1024) and the loss alone, eager."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hardnetnas_b200.hardnet import HardNet
from hardnetnas_b200.losses import loss_HardNet
from bench import timeit
torch.manual_seed(0)
m = HardNet().cuda().eval()
a = torch.rand(1024, 1, 32, 32, device="cuda"); p = a + 0.1 * torch.randn_like(a)
da, dp = m(a), m(p)
print("loss only %.4f ms" % timeit(lambda: loss_HardNet(da, dp, anchor_swap=True), 50))
print("step      %.4f ms" % timeit(lambda: loss_HardNet(m(a), m(p), anchor_swap=True), 20))
