#!/usr/bin/env python
"""Headline benchmark of the HardNet hot path on B200 (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of HardNet.forward over this rank's shard of synthetic 32x32 patches
(BASELINE.json configs[2], bulk extraction: 4,194,304 patches over 8 GPUs = 524,288 patches per GPU per step,
weak scaling). Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "descriptor_patches_per_sec"
UNIT = "patches/s"
PATCHES_PER_GPU = 524288          # configs[2] shard at 8 GPUs
FLOP_PER_PATCH = 78184448
GEN_CHUNK = 65536


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--patches-per-gpu", type=int, default=PATCHES_PER_GPU)
    ap.add_argument("--no-extras", action="store_true", help="skip the loss / matching side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": "BASELINE configs[2]: bulk descriptor extraction, HardNet fp32-in/fp32-out forward, "
                    f"{args.patches_per_gpu} synthetic 32x32 patches per GPU per step "
                    f"({args.patches_per_gpu * world} per step over {world} GPU(s); 4,194,304 at 8 GPUs)",
        "patches_per_gpu_per_step": args.patches_per_gpu,
        "weights": "reference init (torch.manual_seed(0), orthogonal gain 0.6), BN running stats randomised (seed 3), eval mode",
        "activations": "fp16 (10-bit mantissa) with fp32 accumulation",
        "l2_policy": "inputs (2 GiB per GPU per step) are larger than L2; no flush needed",
        "parallelism": f"dp{world} (patch shards, no data-path collective)",
    }


# ---------------------------------------------------------------------------------------------------------
# reference arm: the oracle port (CPU restatement of the reference's PyTorch path) on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_throughput(sample: int, repeats: int = 1):
    from oracle import hardnet_oracle, synth
    w, m, v = synth.hardnet_weights_from_seed(0, 3)
    x = synth.make_patches(sample, 1234, edge_cases=False)
    hardnet_oracle.hardnet_forward(x[:256], w, m, v)  # warm the thread pool / oneDNN primitives
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        hardnet_oracle.hardnet_forward(x, w, m, v)
        best = min(best, time.perf_counter() - t0)
    return sample / best, best


def cpu_side_baselines():
    """The oracle port of the other two hot-path pieces on the host cores (bounded samples): loss_HardNet at N = 1024
    (hardnet/Losses.py:87-154) and NN + ratio matching of an 8192-query chunk against 65 536 gallery rows
    (FDLNet-master/utils/eval_utils.py:113-114,168-175; the reference's full-row sort is replaced by top-2)."""
    from oracle import losses_oracle, synth
    a = synth.unit_vectors(1024, 128, 3)
    p = torch.nn.functional.normalize(a + 0.3 * synth.unit_vectors(1024, 128, 4), dim=1)
    losses_oracle.loss_hardnet(a, p, True)
    t0 = time.perf_counter()
    for _ in range(10):
        losses_oracle.loss_hardnet(a, p, True)
    loss_ms = (time.perf_counter() - t0) / 10 * 1e3
    q, g, _ = synth.make_match_set(8192, 65536, seed=11)
    t0 = time.perf_counter()
    losses_oracle.ratio_match(q, g, 0.7)
    match_s = time.perf_counter() - t0
    return {"loss_hardnet_n1024_ms": loss_ms, "match_8192x65536_s": match_s,
            "match_pairs_per_sec": 8192 * 65536 / match_s}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import hardnet_oracle, synth
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    sample = 2048
    w, m, v = synth.hardnet_weights_from_seed(0, 3)
    x = synth.make_patches(sample, 1234, edge_cases=False)
    for _ in range(args.warmup):
        hardnet_oracle.hardnet_forward(x, w, m, v)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hardnet_oracle.hardnet_forward(x, w, m, v)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(world, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} patches per step (bounded sample of the workload), oracle/hardnet_oracle.py "
                                   "= torch CPU fp32 restatement of hardnet/HardNet.py:312-315, pinned to the reference by tests/golden"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, device_index: int):
        self.rows = []
        self.proc = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", uuid], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_start, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, power, reasons = [], [], [], set()
        for t, line in self.rows:
            if t < t_start or t > t_end + 0.2:
                continue
            f = [c.strip() for c in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(self.NAMES, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_device_patches(n, device, first_chunk_id):
    """uniform noise smoothed 5x5, generated on the device chunk by chunk from seed 1000 + global chunk id."""
    out = torch.empty((n, 1, 32, 32), dtype=torch.float32, device=device)
    g = torch.Generator(device=device)
    for i, s in enumerate(range(0, n, GEN_CHUNK)):
        m = min(GEN_CHUNK, n - s)
        g.manual_seed(1000 + first_chunk_id + i)
        x = torch.rand((m, 1, 32, 32), generator=g, device=device)
        out[s:s + m] = torch.nn.functional.avg_pool2d(x, 5, 1, 2)
    return out


def randomize_bn_stats(state_dict, seed=3):
    """running_mean ~ 0.1 N(0,1), running_var ~ U(0.5,1.5) (same recipe and seed as the parity tests)."""
    g = torch.Generator().manual_seed(seed)
    out = dict(state_dict)
    for k in sorted(out.keys()):
        if k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(out[k].shape, generator=g)
        elif k.endswith("running_var"):
            out[k] = 0.5 + torch.rand(out[k].shape, generator=g)
    return out


def load_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops_sustained": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "tflops_burst": d.get("bf16_tflops"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic(stage_name):
    p = REPO / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(stage_name)
        except Exception:
            return None
    return None


def run_b200(args):
    import torch.distributed as dist
    from hardnetnas_b200 import _lib
    from hardnetnas_b200.extract import DescriptorExtractor
    from hardnetnas_b200.hardnet import HardNet

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lib = _lib.load()
    torch.manual_seed(0)
    model = HardNet()
    model.load_state_dict(randomize_bn_stats(model.state_dict(), 3))
    model = model.to(device).eval()

    P = args.patches_per_gpu
    chunks_per_rank = (P + GEN_CHUNK - 1) // GEN_CHUNK
    x = make_device_patches(P, device, rank * chunks_per_rank)
    out = torch.empty((P, 128), dtype=torch.float32, device=device)

    # ---- warm-up; the first warm-up step is instrumented stage by stage to find the dominant kernel ----
    model(x[:4096], out=out[:4096])
    model.profile_enable(0x7F)
    for w in range(max(args.warmup, 3)):
        model(x, out=out)
        if w == 0:
            torch.cuda.synchronize()
            ms_all, n_all = model.profile_read()
            model.profile_enable(0)
    torch.cuda.synchronize()
    dom = max(range(7), key=lambda i: ms_all[i])
    stage_share = {HardNet.STAGE_NAMES[i]: round(ms_all[i] / max(sum(ms_all), 1e-9), 4) for i in range(7)}

    # ---- timed region: device-resident inputs ------------------------------------------------------------
    model.profile_enable(1 << dom)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    launches0 = lib.hn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        model(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    barrier()
    launches = lib.hn_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_start, t_end) if sampler else None
    dom_ms, dom_n = model.profile_read()
    model.profile_enable(0)
    value = world * P * args.steps / (ms_total / 1e3)

    peaks = load_peaks()
    dom_flops = 2.0 * HardNet.STAGE_MACS[dom] * P * args.steps      # algorithmic FLOPs of that stage over the region
    achieved_tflops = dom_flops / (dom_ms[dom] / 1e3) / 1e12
    flops_per_launch = dom_flops / max(dom_n[dom], 1)
    roofline = {
        "kernel": HardNet.STAGE_NAMES[dom], "bound": "tensor", "achieved": achieved_tflops, "peak": peaks["tflops_sustained"],
        "unit": "TFLOP/s", "frac": achieved_tflops / peaks["tflops_sustained"], "traffic": load_traffic(HardNet.STAGE_NAMES[dom]),
        "peak_source": peaks["source"] + ", sustained bf16 dense (kernel timed inside a long step)",
        "launches": dom_n[dom], "avg_launch_ms": dom_ms[dom] / max(dom_n[dom], 1), "algorithmic_flop_per_launch": flops_per_launch,
        "stage_time_share_warmup": stage_share,
        "whole_path": {"achieved": value / world * FLOP_PER_PATCH / 1e12, "unit": "TFLOP/s",
                       "frac": value / world * FLOP_PER_PATCH / 1e12 / peaks["tflops_sustained"]},
    }

    # ---- end to end: pinned host buffers, H2D + forward + D2H inside the timed region ------------------------
    h_in = torch.empty((P, 1, 32, 32), dtype=torch.float32, pin_memory=True)
    h_in.copy_(x)
    h_out = torch.empty((P, 128), dtype=torch.float32, pin_memory=True)
    del x, out
    torch.cuda.empty_cache()
    ext = DescriptorExtractor(model, device=device)
    ext(h_in, h_out)
    e2e_steps = max(1, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ext(h_in, h_out)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = {"value": world * P * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": world * P * 4096,
           "d2h_bytes_per_step": world * P * 512, "steps": e2e_steps,
           "api": f"hardnetnas_b200.extract.DescriptorExtractor (pinned host in/out, {ext.batch}-patch pipelined batches)"}

    extras = {}
    if not args.no_extras:
        # same end-to-end pipeline fed with uint8 patches (what patch datasets store; input_norm makes the scale irrelevant):
        # 1 KB instead of 4 KB per patch over PCIe - shows how much of the e2e gap at N > 1 is host-to-device bytes
        h_u8 = torch.empty((P, 1, 32, 32), dtype=torch.uint8, pin_memory=True)
        h_u8.copy_((h_in * 255.0).round_().clamp_(0, 255))
        ext8 = DescriptorExtractor(model, device=device, in_dtype=torch.uint8)
        ext8(h_u8, h_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ext8(h_u8, h_out)
        torch.cuda.synchronize()
        t_u8 = max_over_ranks(time.perf_counter() - t0)
        barrier()
        del ext8, h_u8
        extras = side_measurements(device, model, rank, world)
        extras["e2e_uint8_input_patches_per_sec"] = world * P * e2e_steps / t_u8

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        v1, t1 = cpu_reference_throughput(1024)
        sample = int(min(32768, max(1024, 1024 * round(10.0 / max(t1, 1e-3)))))
        v, t = cpu_reference_throughput(sample)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{sample} patches ({t:.1f} s), oracle/hardnet_oracle.py (torch CPU fp32 restatement of the reference forward)"}
        cpu_baseline.update(cpu_side_baselines())

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic", "config": workload_config(args, world), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "extra": extras,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def side_measurements(device, model, rank, world):
    """configs[1] (forward + loss_HardNet at batch 1024) latency and configs[3] (64k x 64k matching) throughput."""
    from hardnetnas_b200 import distributed as hd
    from hardnetnas_b200.losses import loss_HardNet
    from hardnetnas_b200.matching import match_top2
    res = {}
    g = torch.Generator(device=device).manual_seed(7 + rank)

    def timeit(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    if rank == 0:
        a = torch.nn.functional.avg_pool2d(torch.rand((1024, 1, 32, 32), generator=g, device=device), 5, 1, 2)
        p = a + 0.1 * torch.randn(a.shape, generator=g, device=device)

        def step():
            da, dp = model(a), model(p)
            return loss_HardNet(da, dp, anchor_swap=True)
        res["config1_forward_loss_batch1024_ms"] = timeit(step, 20)
        # the same step captured once into a CUDA graph (the calls are stream-ordered, no hidden synchronisation)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            res["config1_forward_loss_batch1024_cuda_graph_ms"] = timeit(graph.replay, 50)
            del graph
        except Exception as exc:   # a capture problem must not take the bench line down
            res["config1_forward_loss_batch1024_cuda_graph_ms"] = None
            print(f"[bench] CUDA graph capture of the config-1 step failed: {exc}", file=sys.stderr)
        da, dp = model(a), model(p)
        res["config1_loss_only_ms"] = timeit(lambda: loss_HardNet(da, dp, anchor_swap=True), 50)
    # matching: 65536 x 65536 overall, query rows and gallery rows sharded over the ranks
    n = 65536
    lo, hi = hd.shard_range(n, rank, world)
    gal = torch.randn((hi - lo, 128), generator=g, device=device)
    gal = gal / gal.norm(dim=1, keepdim=True)
    q = gal + 0.04 * torch.randn(gal.shape, generator=g, device=device)
    q = q / q.norm(dim=1, keepdim=True)
    if world == 1:
        ms = timeit(lambda: match_top2(q, gal), 10)
    else:
        import torch.distributed as dist

        g_counts = [hd.shard_range(n, r, world)[1] - hd.shard_range(n, r, world)[0] for r in range(world)]

        def fn():
            hd.match_sharded(q, gal, g_counts=g_counts)
        for _ in range(2):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 10], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        # "existing implementation on the same box" (BASELINE.md section 3): the reference's stock torch modules on this
        # GPU (cuDNN / cuBLAS), fp32 with TF32 allowed and fp16 autocast + channels_last; 65 536-patch batches
        xs = torch.nn.functional.avg_pool2d(torch.rand((65536, 1, 32, 32), generator=g, device=device), 5, 1, 2)
        with torch.no_grad():
            ms_tf32 = timeit(lambda: model.forward_stock(xs), 5)
            res["stock_torch_gpu_tf32_patches_per_sec"] = 65536 / (ms_tf32 / 1e3)
            with torch.autocast("cuda", dtype=torch.float16):
                ms_amp = timeit(lambda: model.forward_stock(xs), 5)
            res["stock_torch_gpu_fp16_autocast_patches_per_sec"] = 65536 / (ms_amp / 1e3)
        del xs
        # configs[4]: NAS-derived descriptor net (wang2) forward at batch 65 536
        from hardnetnas_b200.nas import SampledDescriptorNet
        torch.manual_seed(0)
        nas = SampledDescriptorNet("wang2").to(device).eval()
        xb = torch.nn.functional.avg_pool2d(torch.rand((65536, 1, 32, 32), generator=g, device=device), 5, 1, 2)
        ob = nas(xb)
        nas_ms = timeit(lambda: nas(xb), 5)
        res["config4_nas_wang2_batch65536_ms"] = nas_ms
        res["config4_nas_wang2_patches_per_sec"] = 65536 / (nas_ms / 1e3)
        res["config4_nas_wang2_frac_of_bf16_peak"] = 65536 * 7272448 / (nas_ms / 1e3) / 1e12 / load_peaks()["tflops_sustained"]
        res["config4_nas_wang2_algorithmic_GBps"] = 65536 * 4608 / (nas_ms / 1e3) / 1e9
        # layer-at-a-time NHWC 16-bit activations move ~562 KB/patch: the HBM fraction that implies
        res["config4_nas_wang2_layerwise_hbm_frac"] = 65536 * 562e3 / (nas_ms / 1e3) / 1e9 / load_peaks()["hbm_gbs"]
        # bytes the 16 kernels of one pass actually move (ncu, profiles/roofline_traffic.json; stem + first 1x1 fused)
        moved = load_traffic("nas_wang2_bytes_per_patch")
        if moved:
            res["config4_nas_wang2_moved_GBps"] = 65536 * moved / (nas_ms / 1e3) / 1e9
            res["config4_nas_wang2_moved_hbm_frac"] = res["config4_nas_wang2_moved_GBps"] / load_peaks()["hbm_gbs"]
        del nas, xb, ob
    res["config3_match_65536x65536_ms"] = ms
    res["config3_match_pairs_per_sec"] = n * n / (ms / 1e3)
    res["config3_match_frac_of_bf16_peak"] = n * n * 256 / (ms / 1e3) / 1e12 / (load_peaks()["tflops_sustained"] * world)
    return res


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else any library prints was diverted to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse_args()
    # libraries (NCCL's version banner, torchrun) write to fd 1: keep stdout clean for the single result line
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
