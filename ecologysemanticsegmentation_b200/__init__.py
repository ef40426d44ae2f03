"""B200-native per-pixel loss and Dice-scoring path, drop-in for the loss / metric functions of
hansk0812/EcologySemanticSegmentation (``loss_functions.py``, ``loss_composite.py``, the
``losses_fn`` of ``train_multiclass.py`` and the scoring of ``test_multiclass.py``).

Python/PyTorch is the host side only (device memory, streams, torch.distributed); the arithmetic
runs in hand-written sm_100a CUDA kernels behind the C ABI of ``include/ecoloss.h``.  CUDA tensors
only -- there is no CPU fallback.
"""
from __future__ import annotations

__version__ = "0.1.0"


class _BCEWithLogits:
    """Stand-in for the reference's module-level ``binary_cross_entropy = torch.nn.BCEWithLogitsLoss()``
    (ecology_semantic_segmentation/__init__.py:24).  Called as ``binary_cross_entropy(input, target)``:
    mean over all elements of ``max(x,0) - x*t + log(1+exp(-|x|))``."""

    def __call__(self, input, target):
        from . import ops
        return ops.leaf7(target, input)[1]

    def __repr__(self):
        return "BCEWithLogitsLoss()  # ecologysemanticsegmentation_b200 CUDA kernel"


binary_cross_entropy = _BCEWithLogits()

from . import loss_functions, loss_composite  # noqa: E402
from .loss_composite import LossList, losses_fn, intersection_loss, union_loss  # noqa: E402,F401

__all__ = ["binary_cross_entropy", "loss_functions", "loss_composite", "LossList", "losses_fn",
           "intersection_loss", "union_loss"]
