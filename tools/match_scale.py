"""Where a sharded 64k x 64k matching call spends its time (run under torchrun on N GPUs of one box)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from hardnetnas_b200 import distributed as hd  # noqa: E402
from hardnetnas_b200.matching import match_top2  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n = 65536
lo, hi = hd.shard_range(n, rank, world)
g = torch.Generator(device=dev).manual_seed(11 + rank)
gal = torch.randn((hi - lo, 128), generator=g, device=dev)
gal = gal / gal.norm(dim=1, keepdim=True)
q = gal + 0.04 * torch.randn(gal.shape, generator=g, device=dev)
q = q / q.norm(dim=1, keepdim=True)
counts = [hd.shard_range(n, r, world)[1] - hd.shard_range(n, r, world)[0] for r in range(world)]
full = hd.all_gather_rows(gal, counts)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = {
    "all_gather (known counts)": timeit(lambda: hd.all_gather_rows(gal, counts)),
    "all_gather (count exchange)": timeit(lambda: hd.all_gather_rows(gal)),
    "local matcher, full gallery": timeit(lambda: match_top2(q, full)),
    "match_sharded (known counts)": timeit(lambda: hd.match_sharded(q, gal, g_counts=counts)),
    "match_sharded (count exchange)": timeit(lambda: hd.match_sharded(q, gal)),
}
if rank == 0:
    for k, v in res.items():
        print(f"N={world} {k:34s} {v:8.3f} ms   ({n * n / v / 1e9:8.1f} G pairs/s whole job)")
dist.destroy_process_group()
