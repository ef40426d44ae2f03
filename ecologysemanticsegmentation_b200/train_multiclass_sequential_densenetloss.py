"""Drop-in for the loss callable of ``ecology_semantic_segmentation/train_multiclass_sequential_densenetloss.py``
(:272-362): per-channel leaves without doubling plus, for C>1, the extra leaf
``losses_fn(g[:,1:2] - g[:,2:3], abs(x[:,1:2] - x[:,2:3]))`` added to channel 1 (:285)."""
from __future__ import annotations

import torch

from . import ops


def losses_fn(x, g, composite_set_theory=False, background_weight=0, early_stopped=False, *, group=None):
    CLASS_INDEX = 1
    ops.nat.require_cuda(x, g)
    if g.shape[CLASS_INDEX] > 1:
        per_channel = ops.PairLeaves.apply(g, x, 0.0, 1.0, 0, group, None, "seq")
        # direct optimisation of the target objective: superset 1 minus subset 2 (:283-285)
        extra = ops.leaf7(g[:, 1:2, :, :] - g[:, 2:3, :, :], torch.abs(x[:, 1:2, :, :] - x[:, 2:3, :, :]), 0.0, 1.0,
                          group=group, key="seq_extra")
        return [a + b for a, b in zip(per_channel, extra)]
    return_losses = list(ops.leaf7(x, g, background_weight, scale=1.0, group=group, key="seq1"))
    if composite_set_theory:
        # :304-320 slice channels 1 and 2 of a 1-channel tensor; the reference fails inside BCEWithLogits
        raise ValueError("Target size (%s) must be the same as input size (%s)"
                         % (tuple(g.shape), tuple(x[:, 1:2].shape)))
    return return_losses
