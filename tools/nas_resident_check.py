"""GPU check of the patch-resident NAS segments: parity against the CPU oracle for several plans and throughput at
batch 65 536 (BASELINE config 5). Every net is built under its own environment (the switches are read in hn_create).

    python tools/nas_resident_check.py [--fast]
"""
import json
import os
import sys
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from hardnetnas_b200.nas import SampledDescriptorNet  # noqa: E402
from hardnetnas_b200.nas.fbnet_modeldef import arch_ops  # noqa: E402
from oracle import nas_oracle, synth  # noqa: E402

MIXED = ["ir_k3_e3_se", "ir_k5_s4", "ir_k3_s2_se", "ir_k5_e3", "ir_k3_s4_se", "ir_k3_e1_se"]
ENV_KEYS = ("HN_NAS_RESIDENT", "HN_NAS_CUT_RATIO", "HN_NAS_MINB", "HN_NAS_GMAX", "HN_NAS_SPLIT")


def build(arch, env, **kw):
    for k in ENV_KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    ops = MIXED if arch == "mixed_se" else arch_ops(arch)
    torch.manual_seed(0)
    net = SampledDescriptorNet(ops, **kw)
    net.load_state_dict(synth.randomize_nas_state(net.state_dict(), 4))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    return net.cuda().eval(), ops, sd


def parity(arch, env):
    net, ops, sd = build(arch, env, chunk_patches=64, head_rows=256)
    plan = net.resident_plan()
    out = {"arch": arch, "env": env, "plan": plan}
    x = synth.make_patches(203, 6, edge_cases=False)
    ref, feats = nas_oracle.nas_forward(x, ops, sd, return_features=True)
    got = net(x.cuda()).float().cpu()
    out["desc_max_abs"] = (got - ref).abs().max().item()
    out["desc_min_cos"] = torch.nn.functional.cosine_similarity(got, ref, dim=1).min().item()
    prog = net.compile_program()
    worst = 0.0
    for stage, op_index in enumerate(prog.stage_end):
        g = net.forward_op(x[:37].cuda(), op_index).float().cpu().permute(0, 3, 1, 2)
        r = feats[stage][:37]
        worst = max(worst, (g - r).abs().max().item() / max(r.abs().max().item(), 1e-9))
    out["stage_rel_err_max"] = worst
    out["ok"] = bool(out["desc_max_abs"] <= 1e-3 and out["desc_min_cos"] >= 0.9999 and worst <= 6e-3)
    return out


def timing(arch, env, batch=65536, iters=5):
    net, ops, sd = build(arch, env)
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.nn.functional.avg_pool2d(torch.rand((batch, 1, 32, 32), generator=g, device="cuda"), 5, 1, 2)
    for _ in range(2):
        y = net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        y = net(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"arch": arch, "env": env, "plan": net.resident_plan(), "ms": ms, "patches_per_sec": batch / ms * 1e3,
            "finite": bool(torch.isfinite(y).all())}


def main():
    fast = "--fast" in sys.argv
    plans = [{}, {"HN_NAS_CUT_RATIO": "2"}, {"HN_NAS_MINB": "1"}, {"HN_NAS_GMAX": "1"}]
    ok = True
    for arch in ("wang2", "wang3", "wang4", "mixed_se"):
        for env in (plans[:2] if fast else plans):
            try:
                r = parity(arch, env)
            except Exception as exc:  # keep going: one broken plan must not hide the others
                r = {"arch": arch, "env": env, "error": repr(exc), "ok": False}
            ok &= r["ok"]
            print("PARITY", json.dumps(r), flush=True)
    tplans = [{"HN_NAS_RESIDENT": "0"}, {}, {"HN_NAS_CUT_RATIO": "2"}, {"HN_NAS_MINB": "1"}, {"HN_NAS_CUT_RATIO": "2", "HN_NAS_MINB": "1"},
              {"HN_NAS_CUT_RATIO": "1000"}]
    for arch in ("wang2",) if fast else ("wang2", "wang3", "wang4"):
        for env in tplans:
            try:
                print("TIMING", json.dumps(timing(arch, env)), flush=True)
            except Exception as exc:
                print("TIMING", json.dumps({"arch": arch, "env": env, "error": repr(exc)}), flush=True)
    print("ALL_OK" if ok else "PARITY_FAILED")


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"elapsed {time.time() - t0:.1f} s")
