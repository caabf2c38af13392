"""Small fixed workload for ncu: a few conv-stack chunks + head, loss and matching (run under gpurun)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hardnetnas_b200.hardnet import HardNet  # noqa: E402
from hardnetnas_b200.losses import loss_HardNet  # noqa: E402
from hardnetnas_b200.matching import match_top2  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "forward"
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    if what == "forward":
        model = HardNet().to(dev).eval()
        x = torch.nn.functional.avg_pool2d(torch.rand(592 * 4, 1, 32, 32, device=dev), 5, 1, 2)
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
    elif what == "match":
        from hardnetnas_b200.matching import mutual_nn_ratio
        from oracle import synth
        q, g, _ = synth.make_match_set(65536, 65536, seed=11)     # BASELINE config 4 set (80 % planted, 20 % unmatched)
        q, g = q.to(dev), g.to(dev)
        for _ in range(3):
            mutual_nn_ratio(q, g, 0.7, return_pairs=False)         # GEMM + block maxima, re-rank, claims, column verification
        a, p = q[:1024].contiguous(), g[:1024].contiguous()
        for _ in range(3):
            loss_HardNet(a, p, anchor_swap=True)
        torch.cuda.synchronize()
    print("done", what)


if __name__ == "__main__":
    main()
