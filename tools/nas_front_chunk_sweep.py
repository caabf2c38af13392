"""Sweep HN_NAS_FRONT_CHUNK (patches per front-kernel + first-reader sub-pass) for the NAS nets at batch 65 536."""
import os, sys
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from tools.nas_resident_check import build, timing  # noqa: E402
from oracle import nas_oracle, synth  # noqa: E402

for arch in ("wang2", "wang3", "wang4"):
    for fc in ("0", "888", "1184", "1480", "1776", "2368", "2960", "4736"):
        os.environ["HN_NAS_FRONT_CHUNK"] = fc
        r = timing(arch, {})
        print(f"SWEEP {arch} front_chunk={fc}: {r['ms']:.3f} ms  {r['patches_per_sec'] / 1e6:.2f} M/s", flush=True)
os.environ["HN_NAS_FRONT_CHUNK"] = "1480"
for arch in ("wang2", "wang3", "wang4", "mixed_se"):
    net, ops, sd = build(arch, {})
    x = synth.make_patches(5000, 6, edge_cases=False)
    ref = nas_oracle.nas_forward(x, ops, sd)
    got = net(x.cuda()).float().cpu()
    print("PARITY", arch, "%.2e" % (got - ref).abs().max().item(), "%.7f" % torch.nn.functional.cosine_similarity(got, ref, dim=1).min().item())
