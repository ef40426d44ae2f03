"""Drop-in for ``ecology_semantic_segmentation/utils/subsets_union.py:8-32``
``return_union_sets_descending_order`` (the label union applied before the loss and the prediction un-union
applied after the sigmoid in ``test_multiclass_sequential_densenetloss.py:66``), as one in-place CUDA pass."""
from __future__ import annotations

import torch

from . import _native as nat


def _exclude_mask(exclude_indices, k):
    mask = 0
    for i in exclude_indices:
        i = int(i)
        if 0 <= i < 64:
            mask |= 1 << i
    return mask


def _union_inplace(ann, axis, exclude_indices, reverse):
    nat.require_cuda(ann)
    if not ann.is_contiguous():
        raise ValueError("return_union_sets_descending_order works in place and needs a contiguous tensor")
    k = ann.shape[axis]
    if k > 64:
        raise ValueError(f"at most 64 entries along the union axis (got {k})")
    outer = 1
    for d in ann.shape[:axis]:
        outer *= d
    inner = 1
    for d in ann.shape[axis + 1:]:
        inner *= d
    if ann.numel() == 0:
        return ann
    rc = nat.lib().eco_union_sets(ann.data_ptr(), nat.dtype_code(ann), outer, k, inner, k * inner, inner,
                                  _exclude_mask(exclude_indices, k), int(bool(reverse)), ann.device.index,
                                  nat.current_stream_ptr(ann.device))
    nat.check(rc, "eco_union_sets")
    return ann


def return_union_sets_descending_order(ann, exclude_indices=[0], reverse=False):
    """utils/subsets_union.py:8-32 -- in place along the class dim (dim 1) of ``ann`` [N,C,H,W]; returns ``ann``."""
    return _union_inplace(ann, 1, exclude_indices, reverse)
