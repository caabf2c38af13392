"""world_size-2 gloo tests of the sharding / gather / merge plumbing (the compute kernel is injected: here
the CPU oracle; on the GPU box the fused kernels)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hardnetnas_b200 import distributed as hd
from oracle import losses_oracle, synth


def _oracle_matcher(q, g):
    lab, ia, da, db = losses_oracle.ratio_match(q, g, 0.7, chunk=256)
    d = losses_oracle.distance_matrix_vector_fdl(q, g)
    i2 = torch.topk(d, 2, dim=-1, largest=False)[1][:, 1]
    return da, db, ia.int(), i2.int()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nq, ng, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q, g, _ = synth.make_match_set(nq, ng, seed=17)
        qlo, qhi = hd.shard_range(nq, rank, world)
        glo, ghi = hd.shard_range(ng, rank, world)
        d1, d2, i1, i2 = hd.match_sharded(q[qlo:qhi], g[glo:ghi], matcher=_oracle_matcher)
        # caller-supplied gallery counts (no count exchange) give the same result, also for uneven shards
        counts = [hd.shard_range(ng, r, world)[1] - hd.shard_range(ng, r, world)[0] for r in range(world)]
        e1, e2, j1, j2 = hd.match_sharded(q[qlo:qhi], g[glo:ghi], matcher=_oracle_matcher, g_counts=counts)
        assert torch.allclose(d1, e1, atol=2e-5) and torch.equal(i1, j1) and torch.allclose(d2, e2, atol=2e-5)
        pairs = hd.mutual_nn_sharded(q[qlo:qhi], g[glo:ghi], matcher=_oracle_matcher)
        x = torch.arange(nq * 3, dtype=torch.float32).view(nq, 3)
        gathered = hd.extract_sharded(None, x, gather=True, forward=lambda t: t * 2)
        ret[rank] = (d1, d2, i1, pairs, gathered)
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 1000, 4194304):
        for world in (1, 2, 3, 8):
            spans = [hd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_matching_equals_single_process():
    world, nq, ng = 2, 301, 515   # ragged: uneven shards on both sides
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, nq, ng, ret), nprocs=world, join=True)
    q, g, _ = synth.make_match_set(nq, ng, seed=17)
    rd1, rd2, ri1, _ = _oracle_matcher(q, g)
    d1 = torch.cat([ret[r][0] for r in range(world)])
    d2 = torch.cat([ret[r][1] for r in range(world)])
    i1 = torch.cat([ret[r][2] for r in range(world)])
    # shard-sized GEMMs may block differently from the full one: indices are exact except at near-ties (<= 1e-6)
    dfull = losses_oracle.distance_matrix_vector_fdl(q, g)
    rows = torch.arange(nq)
    assert (dfull[rows, i1.long()] - dfull[rows, ri1.long()]).abs().max().item() <= 1e-6
    # d = sqrt(2 - 2 a.p): a rounding difference of the fp32 dot product (MKL picks other kernels / summation orders for
    # shard-shaped GEMMs) is amplified by 1 / d ~ 2.3 for matched pairs; 2e-5 is far below any plumbing error
    err1, err2 = (d1 - rd1).abs().max().item(), (d2 - rd2).abs().max().item()
    assert err1 <= 2e-5 and err2 <= 2e-5, (err1, err2)
    pairs = torch.cat([ret[r][3] for r in range(world)])
    ref_pairs = losses_oracle.mutual_nn(q, g)
    a = {tuple(r) for r in pairs.tolist()}
    b = {tuple(r) for r in ref_pairs.tolist()}
    assert len(a ^ b) <= 2 and len(a & b) >= len(b) - 2   # identical up to near-tied rows
    x = torch.arange(nq * 3, dtype=torch.float32).view(nq, 3) * 2
    for r in range(world):
        assert torch.equal(ret[r][4], x)
