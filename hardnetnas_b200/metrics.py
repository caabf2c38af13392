"""Evaluation metrics of the reference's test loop on the device:
pairwise descriptor distances (hardnet/HardNet.py:458) and ErrorRateAt95Recall (hardnet/EvalMetrics.py:6-19),
so the distances of a test epoch never make the per-batch device->host trip of HardNet.py:459-461.

CUDA tensors run on the library's own kernels (csrc/metrics.cu): one row-distance kernel, and a radix SELECTION of the
threshold element instead of the reference's full host-side sort (one launch, one 32-byte read back). CPU tensors evaluate
the reference's expression sequence in torch.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def pair_distances(out_a: torch.Tensor, out_p: torch.Tensor) -> torch.Tensor:
    """torch.sqrt(torch.sum((out_a - out_p) ** 2, 1)) — hardnet/HardNet.py:458."""
    if out_a.is_cuda and out_a.dim() == 2 and out_a.size(1) == 128 and out_a.shape == out_p.shape:
        lib = _lib.load()
        a, p = out_a.detach().float().contiguous(), out_p.detach().float().contiguous()
        out = torch.empty(a.size(0), dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            _lib.check(lib.hn_pair_distances(C.c_void_p(a.data_ptr()), C.c_void_p(p.data_ptr()), a.size(0), C.c_void_p(out.data_ptr()),
                                             _stream()), "hn_pair_distances")
        return out
    return torch.sqrt(torch.sum((out_a - out_p) ** 2, 1))


def fpr95_counts(labels: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
    """Device int64[4] = (FP, TN, #positives, threshold_index) of ErrorRateAt95Recall; no host synchronisation."""
    lib = _lib.load()
    s = scores.detach().float().contiguous().view(-1)
    lab = (labels.detach().view(-1) != 0).to(torch.uint8).contiguous()
    assert s.is_cuda and lab.device == s.device and s.numel() == lab.numel() and s.numel() >= 1
    out = torch.empty(4, dtype=torch.int64, device=s.device)
    with torch.cuda.device(s.device):
        _lib.check(lib.hn_fpr95(C.c_void_p(s.data_ptr()), C.c_void_p(lab.data_ptr()), s.numel(), C.c_void_p(out.data_ptr()), _stream()),
                   "hn_fpr95")
    return out


def ErrorRateAt95Recall(labels: torch.Tensor, scores: torch.Tensor) -> float:
    """hardnet/EvalMetrics.py:6-19: false-positive rate at the distance threshold that recalls 95 % of the matching pairs.
    `scores` = 1 / (distance + 1e-8) as in HardNet.py:472."""
    if scores.is_cuda:
        fp, tn = fpr95_counts(labels, scores)[:2].tolist()
        return float(fp) / float(fp + tn)
    distances = 1.0 / (scores + 1e-8)
    recall_point = 0.95
    order = torch.argsort(distances, stable=True)
    lab = labels[order].to(torch.int64)
    csum = torch.cumsum(lab, 0)
    target = recall_point * float(lab.sum().item())
    threshold_index = int(torch.nonzero(csum >= target)[0].item())   # np.argmax of the boolean array = first True
    fp = int((lab[:threshold_index] == 0).sum().item())
    tn = int((lab[threshold_index:] == 0).sum().item())
    return float(fp) / float(fp + tn)
