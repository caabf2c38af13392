"""Sharded forward matching (64k x 64k) under torchrun: comm-inclusive time for NCCL communicators with different CTA limits,
against the compute-only time. Shows how much of the gap is the collective itself and how much is NCCL's kernels taking SMs
from the persistent matching GEMM."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from hardnetnas_b200 import _ops, distributed as hd  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 65536
    q_all, g_all, _ = synth.make_match_set(n, n, seed=11)
    lo, hi = hd.shard_range(n, rank, world)
    q, g = q_all[lo:hi].to(dev), g_all[lo:hi].to(dev)
    cnt = hi - lo

    def timed(fn, iters=20):
        for _ in range(5):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g_full = hd.all_gather_rows(g, [cnt] * world)
    g16 = _ops.pack_descriptors(g_full)
    res = {"compute_only": timed(lambda: _ops.match_top2(q, g_full, g16=g16))}
    side = torch.cuda.Stream()
    for max_ctas in (0, 16, 8, 4, 2):
        if max_ctas:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = max_ctas
            opts.config.min_ctas = 1
            grp = dist.new_group(ranks=list(range(world)), pg_options=opts)
        else:
            grp = dist.group.WORLD

        def fn():
            p_local = _ops.pack_descriptors(g)
            packed = torch.empty((n, 128), dtype=torch.float16, device=dev)
            dist.all_gather_into_tensor(packed, p_local, group=grp)
            rows = torch.empty((n, 128), dtype=torch.float32, device=dev)
            work = dist.all_gather_into_tensor(rows, g, group=grp, async_op=True)
            ready = torch.cuda.Event()
            with torch.cuda.stream(side):
                work.wait()
                ready.record(side)
            return _ops.match_top2(q, rows, g16=packed, g_ready_event=ready)

        def gather_only():
            p_local = _ops.pack_descriptors(g)
            packed = torch.empty((n, 128), dtype=torch.float16, device=dev)
            dist.all_gather_into_tensor(packed, p_local, group=grp)
            return packed
        res[f"max_ctas_{max_ctas or 'default'}"] = {"match": timed(fn), "fp16_gather_only": timed(gather_only)}
    if rank == 0:
        print("MATCH_SCALE", world, res)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
