"""Verbose per-stage parity report for the NAS descriptor nets (run on the GPU box; diagnostic, not a test)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import nas_oracle, synth  # noqa: E402
from hardnetnas_b200.nas import SampledDescriptorNet  # noqa: E402
from hardnetnas_b200.nas.fbnet_modeldef import arch_ops  # noqa: E402

MIXED = ["ir_k3_e3_se", "ir_k5_s4", "ir_k3_s2_se", "ir_k5_e3", "ir_k3_s4_se", "ir_k3_e1_se"]


def main():
    for arch in ("wang2", "wang3", "wang4", "mixed_se"):
        ops = MIXED if arch == "mixed_se" else arch_ops(arch)
        torch.manual_seed(0)
        net = SampledDescriptorNet(ops)
        net.load_state_dict(synth.randomize_nas_state(net.state_dict(), 4))
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        net = net.cuda().eval()
        x = synth.make_patches(32, 1234, edge_cases=False)
        ref, feats = nas_oracle.nas_forward(x, ops, sd, return_features=True)
        prog = net.compile_program()
        print(arch, "ops:", [(o.kind, o.cin, o.cout, o.hin, o.hout, o.src, o.dst, o.res) for o in prog.ops], flush=True)
        for stage, op_index in enumerate(prog.stage_end):
            got = net.forward_op(x.cuda(), op_index).float().cpu().permute(0, 3, 1, 2)
            err = (got - feats[stage]).abs()
            print(f"  stage {stage} op {op_index}: ref|max| {feats[stage].abs().max():.4f} max err {err.max():.3e} mean err {err.mean():.3e}", flush=True)
        got = net(x.cuda()).cpu()
        print(f"  descriptors: max abs {(got - ref).abs().max():.3e} min cos {torch.nn.functional.cosine_similarity(got, ref).min():.7f}")
        xb = synth.make_patches(65536, 5, edge_cases=False).cuda()
        for _ in range(2):
            net(xb)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            net(xb)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"  throughput: {65536 / dt / 1e6:.3f} M patches/s ({dt * 1e3:.2f} ms per 65536)", flush=True)


if __name__ == "__main__":
    main()
