"""Sharded 64k x 64k matching under torchrun: peer-copy exchange (symmetric memory) vs NCCL all_gathers vs compute only."""
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
    counts = [hi - lo] * world

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

    g_full = hd.all_gather_rows(g, counts)
    g16 = _ops.pack_descriptors(g_full)
    ref = _ops.match_top2(q, g_full, g16=g16)
    res = {"compute_only": timed(lambda: _ops.match_top2(q, g_full, g16=g16))}
    for mode, dbg in (("1", ""), ("0", "")):
        os.environ["HN_P2P_GATHER"] = mode
        got = hd.match_sharded(q, g, g_counts=counts)
        ok = all(torch.equal(a, b) for a, b in zip(got, ref))
        res[f"p2p={mode} {dbg}"] = {"match": timed(lambda: hd.match_sharded(q, g, g_counts=counts)),
                              "mutual": timed(lambda: hd.mutual_nn_ratio_sharded(q, g, counts, counts)),
                              "gather_only": timed(lambda: hd._gather_packed_then_rows(g, counts)), "equal": ok,
                              "peer_exchange_failed": hd._PeerExchange._failed}
    if rank == 0:
        print("MATCH_SCALE3", world, res)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
