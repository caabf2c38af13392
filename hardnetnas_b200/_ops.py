"""Tensor-level wrappers over the distance / matching entry points of the C ABI."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_ws_cache: dict = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Scratch reused across calls (grown on demand), so the hot calls never allocate. One buffer per (device, stream):
    calls on different streams may run concurrently and must not share packed operands or partial minima; a buffer that is
    outgrown is handed back to the caching allocator only after the work queued on its stream (record_stream)."""
    stream = torch.cuda.current_stream(device)
    key = (device.type, device.index, stream.cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            ws.record_stream(stream)
        ws = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _require_cuda_f32(name: str, t: torch.Tensor) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.HardnetB200Error(f"{name}: the fused path runs on B200 CUDA tensors only (no CPU fallback)")
    assert t.dim() == 2, "Inputd must be a 2D matrix."
    if t.size(1) != 128:
        raise ValueError(f"{name}: descriptors must be 128-d, got {t.size(1)}")
    return t.detach().float().contiguous()


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def dist_min(a: torch.Tensor, p: torch.Tensor, form: int, loss_mask: bool, swap: bool, a_xy: torch.Tensor | None = None,
             p_xy: torch.Tensor | None = None, nei_c: float = 0.0):
    """Returns dict(pos, row_min, row_arg[, col_min, col_arg]) — see hn_dist_min / hn_dist_min_ex in include/hardnet_b200.h.
    a_xy / p_xy ([N,2] keypoint coordinates) switch on the neighbour mask of HardNetNeiMask.loss."""
    lib = _lib.load()
    a = _require_cuda_f32("dist_min", a)
    p = _require_cuda_f32("dist_min", p)
    na, npos = a.size(0), p.size(0)
    dev = a.device
    nei = a_xy is not None
    if nei:
        a_xy = a_xy.detach().to(dev, torch.float32).contiguous()
        p_xy = p_xy.detach().to(dev, torch.float32).contiguous()
        assert tuple(a_xy.shape) == (na, 2) and tuple(p_xy.shape) == (npos, 2), "keypoint coordinates must be [N,2]"
    with torch.cuda.device(dev):
        ws = _workspace(dev, lib.hn_dist_workspace_bytes(na, npos, 1))
        has_pos = loss_mask or nei
        out = {
            "pos": torch.empty(min(na, npos), dtype=torch.float32, device=dev) if has_pos else None,
            "row_min": torch.empty(na, dtype=torch.float32, device=dev),
            "row_arg": torch.empty(na, dtype=torch.int32, device=dev),
            "col_min": torch.empty(npos, dtype=torch.float32, device=dev) if swap else None,
            "col_arg": torch.empty(npos, dtype=torch.int32, device=dev) if swap else None,
        }
        flags = (_lib.HN_FLAG_LOSS_MASK if loss_mask else 0) | (_lib.HN_FLAG_SWAP if swap else 0) | (_lib.HN_FLAG_NEI_MASK if nei else 0)
        _lib.check(lib.hn_dist_min_ex(_ptr(a), _ptr(p), na, npos, form, flags, _ptr(a_xy), _ptr(p_xy), C.c_float(nei_c),
                                      _ptr(out["pos"]), _ptr(out["row_min"]), _ptr(out["row_arg"]), _ptr(out["col_min"]),
                                      _ptr(out["col_arg"]), _ptr(ws), ws.numel(), _stream_ptr()), "hn_dist_min")
    return out


def loss_hardnet(anchor: torch.Tensor, positive: torch.Tensor, margin: float, anchor_swap: bool) -> torch.Tensor:
    lib = _lib.load()
    a = _require_cuda_f32("loss_HardNet", anchor)
    p = _require_cuda_f32("loss_HardNet", positive)
    n = a.size(0)
    dev = a.device
    with torch.cuda.device(dev):
        ws = _workspace(dev, lib.hn_dist_workspace_bytes(n, n, 1))
        out = torch.empty((), dtype=torch.float32, device=dev)
        _lib.check(lib.hn_loss_hardnet(_ptr(a), _ptr(p), n, C.c_float(margin), int(bool(anchor_swap)), _ptr(out), _ptr(ws),
                                       ws.numel(), _stream_ptr()), "hn_loss_hardnet")
    return out


def pack_descriptors(x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[n,128] fp32 unit rows -> the fp16 operand rows of the matching GEMM (x * 2^8), see hn_pack_descriptors.
    `out`: optional destination (e.g. a symmetric-memory exchange buffer)."""
    lib = _lib.load()
    x = _require_cuda_f32("pack_descriptors", x)
    if out is None:
        out = torch.empty((x.size(0), 128), dtype=torch.float16, device=x.device)
    elif not (out.is_cuda and out.dtype == torch.float16 and tuple(out.shape) == (x.size(0), 128) and out.is_contiguous()):
        raise ValueError("pack_descriptors: out must be a contiguous CUDA fp16 [n,128] tensor")
    if x.size(0):
        with torch.cuda.device(x.device):
            _lib.check(lib.hn_pack_descriptors(_ptr(x), x.size(0), _ptr(out), _stream_ptr()), "hn_pack_descriptors")
    return out


def pack_descriptors_multicast(x: torch.Tensor, mc16_ptr: int, mc32_ptr: int) -> None:
    """Pack x:[n,128] and push the fp16 and the fp32 rows to multicast addresses (hn_pack_descriptors_multicast)."""
    lib = _lib.load()
    x = _require_cuda_f32("pack_descriptors_multicast", x)
    with torch.cuda.device(x.device):
        _lib.check(lib.hn_pack_descriptors_multicast(_ptr(x), x.size(0), C.c_void_p(mc16_ptr), C.c_void_p(mc32_ptr), _stream_ptr()),
                   "hn_pack_descriptors_multicast")


def match_top2(q: torch.Tensor, g: torch.Tensor, g_offset: int = 0, q16: torch.Tensor | None = None,
               g16: torch.Tensor | None = None, g_ready_event: torch.cuda.Event | None = None,
               block_max: torch.Tensor | None = None):
    """(d1, d2, i1, i2): nearest / second-nearest gallery row per query in the FDLNet distance form.
    q16 / g16: operands already packed by `pack_descriptors` (skips the packing kernels); g_ready_event: the exact re-rank
    (the only reader of the fp32 gallery `g`) waits for it, so `g` may still be arriving while the GEMM runs.
    block_max: optional fp32 tensor of `block_max_elems(nq, ng)` elements that receives the GEMM's per-cell maxima (column side
    of mutual NN, see hn_match_mutual)."""
    lib = _lib.load()
    q = _require_cuda_f32("match", q)
    g = _require_cuda_f32("match", g)
    nq, ng = q.size(0), g.size(0)
    for t, n in ((q16, nq), (g16, ng)):
        if t is not None and not (t.is_cuda and t.dtype == torch.float16 and tuple(t.shape) == (n, 128) and t.is_contiguous()):
            raise ValueError("match: packed operands must be contiguous CUDA fp16 [n,128] tensors from pack_descriptors")
    dev = q.device
    with torch.cuda.device(dev):
        ws = _workspace(dev, lib.hn_dist_workspace_bytes(nq, ng, 0))
        d1 = torch.empty(nq, dtype=torch.float32, device=dev)
        d2 = torch.empty(nq, dtype=torch.float32, device=dev)
        i1 = torch.empty(nq, dtype=torch.int32, device=dev)
        i2 = torch.empty(nq, dtype=torch.int32, device=dev)
        ev = C.c_void_p(g_ready_event.cuda_event) if g_ready_event is not None else C.c_void_p(0)
        _lib.check(lib.hn_match_ex(_ptr(q), _ptr(g), _ptr(q16), _ptr(g16), nq, ng, g_offset, _ptr(d1), _ptr(d2), _ptr(i1),
                                   _ptr(i2), _ptr(block_max), _ptr(ws), ws.numel(), ev, _stream_ptr()), "hn_match")
    return d1, d2, i1, i2


def block_max_elems(nq: int, ng: int) -> int:
    return int(_lib.load().hn_block_max_elems(nq, ng))


def match_mutual(q: torch.Tensor, g: torch.Tensor, q16: torch.Tensor | None = None, g16: torch.Tensor | None = None):
    """(d1, d2, i1, i2, mutual bool[Nq]) from ONE matching GEMM: hn_match_mutual."""
    lib = _lib.load()
    q = _require_cuda_f32("match_mutual", q)
    g = _require_cuda_f32("match_mutual", g)
    nq, ng = q.size(0), g.size(0)
    dev = q.device
    with torch.cuda.device(dev):
        ws = _workspace(dev, lib.hn_mutual_workspace_bytes(nq, ng))
        d1 = torch.empty(nq, dtype=torch.float32, device=dev)
        d2 = torch.empty(nq, dtype=torch.float32, device=dev)
        i1 = torch.empty(nq, dtype=torch.int32, device=dev)
        i2 = torch.empty(nq, dtype=torch.int32, device=dev)
        mutual = torch.empty(nq, dtype=torch.bool, device=dev)
        _lib.check(lib.hn_match_mutual(_ptr(q), _ptr(g), _ptr(q16), _ptr(g16), nq, ng, _ptr(d1), _ptr(d2), _ptr(i1), _ptr(i2),
                                       _ptr(mutual), _ptr(ws), ws.numel(), _stream_ptr()), "hn_match_mutual")
    return d1, d2, i1, i2, mutual


def mutual_claims(i1: torch.Tensor, d1: torch.Tensor, d2: torch.Tensor, q_offset: int, claim: torch.Tensor) -> torch.Tensor:
    """atomicMin of (distance, global row) into claim[int64, Ng] (pre-filled with INT64_MAX = unclaimed); returns the per
    32-row-block minimum of d2 the verification needs."""
    lib = _lib.load()
    rbmin = torch.empty((i1.numel() + 31) // 32, dtype=torch.float32, device=i1.device)
    with torch.cuda.device(i1.device):
        _lib.check(lib.hn_mutual_claims(_ptr(i1), _ptr(d1), _ptr(d2), i1.numel(), q_offset, _ptr(claim), claim.numel(), _ptr(rbmin),
                                        _stream_ptr()), "hn_mutual_claims")
    return rbmin


def mutual_verify(q: torch.Tensor, q_offset: int, g: torch.Tensor, claim: torch.Tensor, block_max: torch.Tensor,
                  rb_min_d2: torch.Tensor, beaten: torch.Tensor):
    """beaten[uint8, Ng] |= some local query row is nearer to g_j than the claimant of column j."""
    lib = _lib.load()
    with torch.cuda.device(q.device):
        _lib.check(lib.hn_mutual_verify(_ptr(q), q.size(0), q_offset, _ptr(g), g.size(0), _ptr(claim), _ptr(block_max),
                                        _ptr(rb_min_d2), _ptr(beaten), _stream_ptr()), "hn_mutual_verify")
