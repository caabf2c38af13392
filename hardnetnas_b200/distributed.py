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


def match_sharded(q_local: torch.Tensor, g_local: torch.Tensor, matcher: Callable | None = None,
                  g_counts: list[int] | None = None):
    """(d1, d2, i1, i2) for this rank's query rows against the gallery of ALL ranks; indices are global
    gallery rows in rank order. `g_counts` = gallery rows held by every rank when the caller knows them (e.g. from
    `shard_range`): it saves the small count exchange and its host synchronisation, which is a visible share of a
    sub-millisecond matching call."""
    matcher = matcher or _default_matcher()
    g_full = all_gather_rows(g_local, g_counts)
    return matcher(q_local, g_full)


def mutual_nn_sharded(q_local: torch.Tensor, g_local: torch.Tensor, matcher: Callable | None = None) -> torch.Tensor:
    """Mutual-NN pairs (global query row, global gallery row) whose query row lives on this rank."""
    matcher = matcher or _default_matcher()
    rank, world = _world()
    g_full = all_gather_rows(g_local)
    q_full = all_gather_rows(q_local)
    fwd_local = matcher(q_local, g_full)[2].long()   # nearest gallery row of my queries
    bwd_local = matcher(g_local, q_full)[2].long()   # nearest query row of my gallery rows
    bwd_full = all_gather_rows(bwd_local)
    q_counts = all_gather_rows(torch.tensor([q_local.size(0)], dtype=torch.long, device=q_local.device))
    q_off = int(q_counts[:rank].sum().item()) if world > 1 else 0
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
