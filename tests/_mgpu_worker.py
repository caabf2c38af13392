"""Worker of tests/test_multigpu_gpu.py: one process per GPU under torchrun (NCCL). Every rank checks the sharded calls of
hardnetnas_b200.distributed against the same problem solved on its own GPU alone; exit code 0 = all checks passed."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from hardnetnas_b200 import distributed as hd  # noqa: E402
from hardnetnas_b200.hardnet import HardNet  # noqa: E402
from hardnetnas_b200.matching import match_top2, mutual_nearest_neighbors  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for nq, ng in ((4096 * world, 8192 * world), (4096 * world - 3, 8192 * world + 5)):   # equal and ragged shards
        q, g, _ = synth.make_match_set(nq, ng, seed=11)
        q, g = q.to(dev), g.to(dev)
        qlo, qhi = hd.shard_range(nq, rank, world)
        glo, ghi = hd.shard_range(ng, rank, world)
        g_counts = [hd.shard_range(ng, r, world)[1] - hd.shard_range(ng, r, world)[0] for r in range(world)]
        q_counts = [hd.shard_range(nq, r, world)[1] - hd.shard_range(nq, r, world)[0] for r in range(world)]
        ref = match_top2(q, g)
        for counts in (g_counts, None):
            got = hd.match_sharded(q[qlo:qhi], g[glo:ghi], g_counts=counts)
            for a, b, name in zip(got, ref, ("d1", "d2", "i1", "i2")):
                assert torch.equal(a, b[qlo:qhi]), f"match_sharded {name} differs (nq={nq}, counts={counts is not None})"
        ref_pairs = mutual_nearest_neighbors(q, g)
        mine = ref_pairs[(ref_pairs[:, 0] >= qlo) & (ref_pairs[:, 0] < qhi)]
        for qc, gc in ((q_counts, g_counts), (None, None)):
            got = hd.mutual_nn_sharded(q[qlo:qhi], g[glo:ghi], q_counts=qc, g_counts=gc)
            assert torch.equal(got, mine), f"mutual_nn_sharded differs (nq={nq})"
    torch.manual_seed(0)
    model = HardNet()
    model.load_state_dict(synth.randomize_bn_stats(model.state_dict(), 3))
    model = model.to(dev).eval()
    x = synth.make_patches(1001, 21).to(dev)
    full = model(x)
    got = hd.extract_sharded(model, x, gather=True)
    assert torch.equal(got, full), "extract_sharded(gather=True) differs from the single-GPU forward"
    lo, hi = hd.shard_range(1001, rank, world)
    assert torch.equal(hd.extract_sharded(model, x), full[lo:hi])
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
