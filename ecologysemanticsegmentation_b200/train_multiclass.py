"""Drop-in for the loss callable of ``ecology_semantic_segmentation/train_multiclass.py``.

``train(net, traindataloader, valdataloader, losses_fn, optimizer, ...)`` (train_multiclass.py:48) receives
the loss as an injected callable (passed at :390) and unpacks exactly 7 values from
``losses_fn(outputs, labels, composite_set_theory=False, background_weight=bg, early_stopped=...)`` (:139-141).
This module provides that callable; the training loop, optimiser and checkpointing stay the reference's.
"""
from __future__ import annotations

from . import ops
from .loss_functions import (binary_cross_entropy_list, classification_dice_list, cross_entropy_list,
                             focal_list)


def return_union_sets_descending_order(ann, exclude_indices=[0]):
    """train_multiclass.py:32-45 -- the twin ``train()`` calls on every label batch (:110).  It walks
    ``ann.shape[0]``, i.e. the BATCH dimension of the [N,C,H,W] labels (the class-dim version lives in
    utils/subsets_union.py); kept as is, in place, as one CUDA pass."""
    from .subsets_union import _union_inplace
    return _union_inplace(ann, 0, exclude_indices, False)


def losses_fn(x, g, composite_set_theory=False, background_weight=0, early_stopped=False, *, group=None,
              from_logits=False):
    """train_multiclass.py:253-303: like ``loss_composite.losses_fn`` but WITHOUT the doubling (:274), returning
    a plain list.  For C>1 the composite flag is ignored (early return :260-262, background_weight dropped);
    for C == 1 the composite branch of the reference raises, and so does this."""
    CLASS_INDEX = 1
    if not isinstance(x, list):
        ops.nat.require_cuda(x, g)
    if not isinstance(g, list) and g.shape[CLASS_INDEX] > 1:
        flags = ops.nat.FLAG_B_LOGIT if from_logits else 0
        return list(ops.PairLeaves.apply(g, x, 0.0, 1.0, flags, group, None, "tm"))

    if isinstance(x, list):
        # deep-supervision branch (:264-267): binary_cross_entropy_list works, the next helper raises
        # TypeError in the reference as well
        bce_loss = binary_cross_entropy_list(x, g)
        ce_loss, fl_loss = cross_entropy_list(x, g), focal_list(x, g, factor=1e-5)
        dice, generalized_dice, twersky_dice, focal_dice = classification_dice_list(x, g, factor=10)
        return [ce_loss, bce_loss, fl_loss, dice, generalized_dice, twersky_dice, focal_dice]

    flags = ops.nat.FLAG_A_LOGIT if from_logits else 0
    return_losses = list(ops.leaf7(x, g, background_weight, scale=1.0, flags=flags, group=group, key="tm1"))
    if composite_set_theory:
        # :276-301 slices channels 1 and 2 of a 1-channel tensor and zips three lists into two names
        raise ValueError("too many values to unpack (expected 2)")
    return return_losses
