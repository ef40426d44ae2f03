"""Exact-integer thresholded Dice counter.  TEST INFRASTRUCTURE.

The reference's own ``torch.sum`` over thresholded fp32 tensors is inexact above 2**24 elements
(SURVEY.md 8(c) "Exact-count caveat"), so the oracle runs the reference's thresholding rule
(ess/test_multiclass.py:68-69) and then sums in int64.
"""
from __future__ import annotations

import numpy as np
import torch

from .torch_port import threshold_inplace

EPS = 1e-7


def batch_counts(logits: torch.Tensor, labels: torch.Tensor, threshold: float):
    """int64 [C,3] = per class (intersection = sum out*lab, |out| = sum out, |lab| = sum lab) after
    ``out = sigmoid(logits); out[out>T]=1; out[out!=1]=0`` -- evaluated on the tensors' own device so the
    sigmoid bits are that device's."""
    out = threshold_inplace(torch.sigmoid(logits), threshold)
    C = labels.shape[1]
    res = np.zeros((C, 3), dtype=np.int64)
    for c in range(C):
        o = out[:, c].to(torch.int64)
        l = labels[:, c].to(torch.int64)
        res[c] = (int((o * l).sum()), int(o.sum()), int(l.sum()))
    return res


def dice_from_counts(counts: np.ndarray):
    """(2 I + eps) / (|out| + |lab| + eps) per class, float64 (test_multiclass.py:80 with loss_functions.py:55-57;
    lab**2 == lab for binary labels)."""
    c = counts.astype(np.float64)
    return (2 * c[:, 0] + EPS) / (c[:, 1] + c[:, 2] + EPS)


def batch_soft_sums(logits: torch.Tensor, labels: torch.Tensor):
    """float64 [C,3] = (sum out*lab, sum out, sum lab**2) for the un-thresholded live path (:80-81)."""
    out = torch.sigmoid(logits).double()
    lab = labels.double()
    C = labels.shape[1]
    return np.array([[float((out[:, c] * lab[:, c]).sum()), float(out[:, c].sum()), float((lab[:, c] ** 2).sum())]
                     for c in range(C)])


def masks_u8(t: torch.Tensor) -> np.ndarray:
    """ess/test_multiclass.py:90-92 (and ess/test_video.py:129-130): ``(t.numpy() * 255).astype(np.uint8)`` on the
    host copy of a float32 tensor (outputs after the sigmoid / threshold rule, labels, images)."""
    return (t.detach().cpu().numpy() * 255).astype(np.uint8)
