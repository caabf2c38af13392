"""Verbose per-stage parity report for the HardNet path (run on the GPU box; diagnostic, not a test)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import hardnet_oracle, synth  # noqa: E402
from hardnetnas_b200.hardnet import HardNet  # noqa: E402


def main():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    for bn_seed in (3, None):
        w, m, v = synth.hardnet_weights_from_seed(0, bn_seed)
        torch.manual_seed(0)
        model = HardNet()
        sd = model.state_dict()
        for i, bi in enumerate(synth.BN_IDX):
            sd[f"features.{bi}.running_mean"] = m[i]
            sd[f"features.{bi}.running_var"] = v[i]
        model.load_state_dict(sd)
        model = model.cuda().eval()
        x = synth.make_patches(64, 1234)
        acts = hardnet_oracle.hardnet_stages(x, w, m, v, upto=6)
        xg = x.cuda()
        for layer in range(1, 7):
            got = model.forward_stage(xg, layer).float().cpu().permute(0, 3, 1, 2)
            ref = acts[layer - 1]
            err = (got - ref).abs()
            print(f"bn={bn_seed} stage {layer}: shape {tuple(got.shape)} ref|max| {ref.abs().max():.4f} "
                  f"max err {err.max():.3e} mean err {err.mean():.3e} got|max| {got.abs().max():.4f}", flush=True)
            if err.max() > 1e-2 * ref.abs().max():
                bad = (err > 1e-2 * ref.abs().max()).nonzero()
                print("   first bad idx (n,c,y,x):", bad[:8].tolist(), "count", len(bad))
        desc = model(xg)
        torch.cuda.synchronize()
        ref = hardnet_oracle.hardnet_forward(x, w, m, v)
        d = desc.cpu()
        print(f"bn={bn_seed} descriptors: max abs {(d - ref).abs().max():.3e} min cos "
              f"{torch.nn.functional.cosine_similarity(d[:-1], ref[:-1]).min():.7f}")
    # quick throughput probe
    x = synth.make_patches(16384, 5, edge_cases=False).cuda()
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        model(x)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"throughput probe: {x.size(0) / dt / 1e6:.3f} M patches/s ({dt * 1e3:.2f} ms per 16384)")


if __name__ == "__main__":
    main()
