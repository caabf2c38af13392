"""Evaluation metrics of the reference's test loop on the device:
pairwise descriptor distances (hardnet/HardNet.py:458) and ErrorRateAt95Recall (hardnet/EvalMetrics.py:6-19),
so the distances of a test epoch never make the per-batch device->host trip of HardNet.py:459-461.
Plain torch ops on CUDA tensors (sort / cumsum): host plumbing around the descriptor kernels, not a kernel itself.
"""
from __future__ import annotations

import torch


def pair_distances(out_a: torch.Tensor, out_p: torch.Tensor) -> torch.Tensor:
    """torch.sqrt(torch.sum((out_a - out_p) ** 2, 1)) — hardnet/HardNet.py:458."""
    return torch.sqrt(torch.sum((out_a - out_p) ** 2, 1))


def ErrorRateAt95Recall(labels: torch.Tensor, scores: torch.Tensor) -> float:
    """hardnet/EvalMetrics.py:6-19 on device tensors: false-positive rate at the distance threshold that recalls
    95 % of the matching pairs. `scores` = 1 / (distance + 1e-8) as in HardNet.py:472."""
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
