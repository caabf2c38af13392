"""One-process-per-GPU sharding of the hot path over torch.distributed (NCCL on the GPU box, gloo in the
CPU tests). The reference has no distributed code; SURVEY.md §8(e) defines the sharding:

  * extraction: patches are independent -> contiguous 1/R split per rank, no data-path collective;
  * matching:   query rows are sharded; every rank holds Ng/R gallery rows and ONE all_gather over NVLink
                gives it the whole gallery, after which its rows' results are final locally. For mutual NN
                the reverse direction is sharded over gallery rows (queries all-gathered) and only the two
                index vectors are exchanged.

The compute kernels are injected (`matcher`), so the plumbing is testable on CPU with gloo.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first n % world ranks get one extra row."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_rows(local: torch.Tensor, counts: list[int] | None = None) -> torch.Tensor:
    """Concatenate row blocks of every rank (variable row counts allowed)."""
    rank, world = _world()
    if world == 1:
        return local
    if counts is None:
        c = torch.tensor([local.size(0)], dtype=torch.long, device=local.device)
        cl = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(cl, c)
        counts = [int(t.item()) for t in cl]
    if len(set(counts)) == 1:
        out = torch.empty((world * counts[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.size(0)] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def _default_matcher():
    from . import _ops
    return _ops.match_top2


def _counts_or_exchange(local_rows: int, counts: list[int] | None, device) -> list[int]:
    rank, world = _world()
    if counts is not None or world == 1:
        return counts if counts is not None else [local_rows]
    c = torch.tensor([local_rows], dtype=torch.long, device=device)
    cl = [torch.zeros_like(c) for _ in range(world)]
    dist.all_gather(cl, c)
    return [int(t.item()) for t in cl]


def _gather_packed_then_rows(local: torch.Tensor, counts: list[int]):
    """Gallery exchange of the B200 path. Returns (rows_fp32, rows_fp16, ready_event): the packed fp16 rows (what the GEMM
    reads: half the bytes, each rank packs only its shard) are complete on the current stream, the fp32 rows (what the exact
    re-rank reads) when `ready_event` fires - they arrive behind the GEMM.
    Preferred transport: PEER COPIES over NVLink out of symmetric-memory send buffers (`_PeerExchange`): the copy engines
    move the shards, no SM is taken from the persistent matching GEMM (NCCL's gather kernels running next to it cost more
    than the transfer itself at 8 GPUs). Fallback: two NCCL all_gathers, the second one asynchronous."""
    from . import _ops
    world = len(counts)
    xchg = _PeerExchange.get(local.device, counts[0]) if _peer_exchange_enabled() else None
    if xchg is not None:
        return xchg.gather(local)
    packed_local = _ops.pack_descriptors(local)
    packed = torch.empty((world * counts[0], 128), dtype=torch.float16, device=local.device)
    dist.all_gather_into_tensor(packed, packed_local)
    rows = torch.empty((world * counts[0], 128), dtype=torch.float32, device=local.device)
    work = dist.all_gather_into_tensor(rows, local.contiguous(), async_op=True)
    side = _side_stream(local.device)
    ready = torch.cuda.Event()
    with torch.cuda.stream(side):
        work.wait()            # the side stream (not the compute stream) waits for the collective
        ready.record(side)
    return rows, packed, ready


def _peer_exchange_enabled() -> bool:
    import os
    return os.environ.get("HN_P2P_GATHER", "1") != "0"


class _PeerExchange:
    """Gallery exchange through symmetric memory (torch.distributed._symmetric_memory: CUDA VMM allocations mapped into every
    rank of the group, with an NVSwitch MULTICAST mapping on NVLS-capable systems). Per call every rank packs its shard and
    PUSHES it - the fp16 operand rows and the fp32 rows - to the multicast address of its slice of the symmetric gallery
    buffers (`hn_pack_descriptors_multicast`, multimem.st): the switch replicates the stores into every GPU's copy, so one
    device-side barrier (~7 us) later each GPU holds the whole gallery. No gather kernel shares the SMs with the persistent
    matching GEMM, no copy engine is programmed 2(R-1) times, the data crosses each GPU's NVLink once.
    The buffers are double-buffered by call parity: a rank passes barrier k + 1 only after every rank has finished reading
    the buffers of call k (their barrier sits behind their re-rank in stream order), so the pushes of call k + 2 are safe."""
    _cache: dict = {}
    _failed = False

    @classmethod
    def get(cls, device, rows: int):
        if cls._failed:
            return None
        key = (device.index, rows)
        if key not in cls._cache:
            try:
                cls._cache[key] = cls(device, rows)
            except Exception as exc:   # no symmetric memory / multicast on this system or build: NCCL path
                import sys
                print(f"[hardnetnas_b200] multicast exchange unavailable ({exc!r}); using NCCL all_gather", file=sys.stderr)
                cls._failed = True
                return None
        return cls._cache[key]

    def __init__(self, device, rows: int):
        import torch.distributed._symmetric_memory as symm_mem
        self.rank, self.world = _world()
        self.rows, self.device = rows, device
        group = dist.group.WORLD
        n = self.world * rows
        self.full16 = [symm_mem.empty((n, 128), dtype=torch.float16, device=device) for _ in range(2)]
        self.full32 = [symm_mem.empty((n, 128), dtype=torch.float32, device=device) for _ in range(2)]
        self.h16 = [symm_mem.rendezvous(t, group) for t in self.full16]
        self.h32 = [symm_mem.rendezvous(t, group) for t in self.full32]
        if any(int(h.multicast_ptr) == 0 for h in self.h16 + self.h32):
            raise RuntimeError("no multicast mapping (NVLS) for the symmetric buffers")
        self.mc16 = [int(h.multicast_ptr) + self.rank * rows * 128 * 2 for h in self.h16]
        self.mc32 = [int(h.multicast_ptr) + self.rank * rows * 128 * 4 for h in self.h32]
        self.calls = 0

    def gather(self, local: torch.Tensor):
        from . import _ops
        par = self.calls & 1
        self.calls += 1
        # One kernel pushes both forms, one barrier publishes them. Measured alternatives at 8 GPUs (tools/match_scale3.py,
        # 64k x 64k, compute-only 0.153 ms): this 0.239 ms; fp16 first and the fp32 rows on a side stream behind the GEMM
        # 0.253 ms (whatever runs next to the persistent GEMM slows it by more than the transfer it hides); peer copies by the
        # copy engines 0.35 ms (2 R small copies per call); two NCCL all_gathers 0.276 ms.
        _ops.pack_descriptors_multicast(local, self.mc16[par], self.mc32[par])
        self.h16[par].barrier(channel=0)      # every rank's pushes of this call have landed (and call k - 2 is fully read)
        return self.full32[par], self.full16[par], None


_side_streams: dict = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def match_sharded(q_local: torch.Tensor, g_local: torch.Tensor, matcher: Callable | None = None,
                  g_counts: list[int] | None = None):
    """(d1, d2, i1, i2) for this rank's query rows against the gallery of ALL ranks; indices are global
    gallery rows in rank order. `g_counts` = gallery rows held by every rank when the caller knows them (e.g. from
    `shard_range`): it saves the small count exchange and its host synchronisation, which is a visible share of a
    sub-millisecond matching call.
    With the default (B200) matcher and equal shards the packed gallery is gathered first and the fp32 rows arrive behind the
    GEMM (`_gather_packed_then_rows`); an injected matcher (CPU tests) or ragged shards use one plain gather."""
    rank, world = _world()
    if matcher is None and world > 1 and q_local.is_cuda:
        counts = _counts_or_exchange(g_local.size(0), g_counts, g_local.device)
        if len(set(counts)) == 1:
            from . import _ops
            g_full, g16_full, ready = _gather_packed_then_rows(g_local, counts)
            return _ops.match_top2(q_local, g_full, g16=g16_full, g_ready_event=ready)
        g_counts = counts
    matcher = matcher or _default_matcher()
    g_full = all_gather_rows(g_local, g_counts)
    return matcher(q_local, g_full)


def mutual_nn_ratio_sharded(q_local: torch.Tensor, g_local: torch.Tensor, q_counts: list[int], g_counts: list[int],
                            threshold: float = 0.7):
    """BASELINE config 4 on R GPUs, as masks over this rank's query rows (no host synchronisation):
    (mutual bool[nq], ratio_label bool[nq], Ia int64[nq] global gallery rows, Da, Db). Equal gallery shards required.
    ONE GEMM per rank (my query rows x the whole gallery, packed gallery gathered first, fp32 rows behind the GEMM). Column
    side: every rank claims the nearest columns of its rows, all_reduce(MIN) of the packed (distance, row) claims (8 B per
    gallery row), every rank checks the global claims against its own rows with the block maxima of its GEMM,
    all_reduce(MAX) of the 1-byte "beaten" flags (hn_mutual_claims / hn_mutual_verify)."""
    from . import _ops
    rank, world = _world()
    dev = q_local.device
    g_full, g16_full, g_ready = _gather_packed_then_rows(g_local, g_counts)
    ng, nq = g_full.size(0), q_local.size(0)
    q_off = sum(q_counts[:rank])
    bm = torch.empty(_ops.block_max_elems(nq, ng), dtype=torch.float32, device=dev)
    d1, d2, fwd, _ = _ops.match_top2(q_local, g_full, g16=g16_full, g_ready_event=g_ready, block_max=bm)
    claim = torch.full((ng,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)   # unclaimed
    rbmin = _ops.mutual_claims(fwd, d1, d2, q_off, claim)
    dist.all_reduce(claim, op=dist.ReduceOp.MIN)     # distances are positive: signed order = unsigned order of the packing
    beaten = torch.zeros(ng, dtype=torch.uint8, device=dev)
    _ops.mutual_verify(q_local, q_off, g_full, claim, bm, rbmin, beaten)
    dist.all_reduce(beaten, op=dist.ReduceOp.MAX)
    fwd = fwd.long()
    i = torch.arange(nq, device=dev) + q_off
    mutual = ((claim[fwd] & 0xFFFFFFFF) == i) & (beaten[fwd] == 0)
    return mutual, (d1 / d2).lt(threshold), fwd, d1, d2


def mutual_nn_sharded(q_local: torch.Tensor, g_local: torch.Tensor, matcher: Callable | None = None,
                      q_counts: list[int] | None = None, g_counts: list[int] | None = None) -> torch.Tensor:
    """Mutual-NN pairs (global query row, global gallery row) whose query row lives on this rank.
    B200 path: one GEMM per rank plus two small all_reduces (see the body). Injected matcher (CPU tests) / ragged gallery
    shards: forward direction over the gathered gallery, backward direction over the gathered queries, backward index vector
    exchanged. With `q_counts` / `g_counts` given no count exchange happens."""
    rank, world = _world()
    qc = _counts_or_exchange(q_local.size(0), q_counts, q_local.device)
    gc = _counts_or_exchange(g_local.size(0), g_counts, g_local.device)
    if matcher is None and world > 1 and q_local.is_cuda and len(set(gc)) == 1:
        mutual, _, fwd, _, _ = mutual_nn_ratio_sharded(q_local, g_local, qc, gc)
        i = torch.arange(q_local.size(0), device=q_local.device) + sum(qc[:rank])
        return torch.stack([i[mutual], fwd[mutual]], dim=1)
    matcher = matcher or _default_matcher()
    g_full = all_gather_rows(g_local, gc)
    q_full = all_gather_rows(q_local, qc)
    fwd_local = matcher(q_local, g_full)[2].long()   # nearest gallery row of my queries
    bwd_local = matcher(g_local, q_full)[2]          # nearest query row of my gallery rows
    bwd_full = all_gather_rows(bwd_local.contiguous(), gc).long()
    q_off = sum(qc[:rank])
    i = torch.arange(q_local.size(0), device=q_local.device) + q_off
    keep = bwd_full[fwd_local] == i
    return torch.stack([i[keep], fwd_local[keep]], dim=1)


def extract_sharded(model, patches: torch.Tensor, gather: bool = False, forward: Callable | None = None) -> torch.Tensor:
    """Each rank runs the forward on its contiguous shard of `patches` (every rank passes the same tensor, or
    a rank may pass only its shard with world==1 semantics). With gather=True all ranks return all
    descriptors, otherwise the local shard's."""
    rank, world = _world()
    lo, hi = shard_range(patches.size(0), rank, world)
    fwd = forward or model
    local = fwd(patches[lo:hi])
    if gather and world > 1:
        counts = [shard_range(patches.size(0), r, world)[1] - shard_range(patches.size(0), r, world)[0] for r in range(world)]
        return all_gather_rows(local, counts)
    return local
