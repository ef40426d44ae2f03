"""Drop-in for the scoring of ``ecology_semantic_segmentation/test_multiclass_sequential_densenetloss.py``.

The sequential model predicts nested unions (whole body, ventral + dorsal, dorsal); its ``test()`` (:31-140) differs from
``test_multiclass.test`` in one line: after ``F.sigmoid(net(x))`` (:62) the predictions are un-unioned,
``return_union_sets_descending_order(test_outputs, reverse=True)`` (:66 -> utils/subsets_union.py:21-27,
``p_c <- |p_c - p_{c+1}|`` for c = C-2 .. 1), before the per-class soft Dice (:97-99; the beam-search list of :84 is empty
there, so no threshold is ever applied).  Here sigmoid, un-union and the Dice sums are ONE read of logits and labels: the
scoring kernel takes the difference in registers (``ECO_EVAL_UNUNION``), no in-place sweep over the predictions.
"""
from __future__ import annotations

from . import test_multiclass as _tm

ORGANS = _tm.ORGANS
DiceAccumulator = _tm.DiceAccumulator


def score_batch(logits, labels, *, group=None, inputs_are_probs=False, return_counts=False):
    """Per-class soft Dice of ``return_union_sets_descending_order(sigmoid(logits), reverse=True)`` against ``labels``,
    float32 [C] on the device (``logits`` are not modified)."""
    return _tm.score_batch(logits, labels, None, group=group, inputs_are_probs=inputs_are_probs,
                           return_counts=return_counts, ununion=True)


def test(net, dataloader, models_dir="models/vgg", results_dir="test_results/", batch_size=1, saved_epoch=-1,
         single_model=False):
    """Same call signature and return value as the reference ``test()`` (:31): mean per-batch soft Dice per organ of the
    un-unioned predictions, a float32 CPU tensor [C] (None when this epoch's results directory already exists)."""
    return _tm.test(net, dataloader, models_dir=models_dir, results_dir=results_dir, batch_size=batch_size,
                    saved_epoch=saved_epoch, single_model=single_model, threshold=None, ununion=True)
