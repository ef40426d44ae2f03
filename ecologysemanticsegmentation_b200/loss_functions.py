"""Drop-in for ``ecology_semantic_segmentation/loss_functions.py``: same names, positional order,
defaults and return conventions, computed by the sm_100a pair-leaf kernels.

Every function takes ``(gt, pred, ...)`` exactly like the reference; internally ``gt`` is slot a and
``pred`` slot b.  Each returns a 0-d CUDA tensor that participates in autograd w.r.t. BOTH arguments.
Quirks of the reference are preserved on purpose.
"""
from __future__ import annotations

import torch

from . import binary_cross_entropy
from . import ops

_M = ops.M_DICE


def _shape(focal_gamma=1.5, alpha=0.5, beta=0.3, focal_dice_gamma=1.8):
    """Keyword parameters of the primitives -> the kernels' shape tuple (None = the reference's defaults, which keep
    the sqrt closed forms; any other value takes the powf instantiation of the same kernels)."""
    t = (float(focal_gamma), float(alpha), float(beta), float(focal_dice_gamma))
    return None if t == ops.nat.DEFAULT_SHAPE else t


def binary_cross_entropy_list(gt, pred):
    """loss_functions.py:13-20 -- sum of up to 6 BCE terms through a CPU buffer (returns a CPU tensor)."""
    sum_arr = torch.zeros(6)
    for idx, (y, p) in enumerate(zip(gt, pred)):
        sum_arr[idx] = cross_entropy_loss(y, p, bce=True)
    return torch.sum(sum_arr)


# loss_functions.py:22-24 -- these three are broken in the reference (torch.sum of a python list, and
# keyword arguments the callees do not take): calling them raises TypeError there, and here.
cross_entropy_list = lambda xL, yL: torch.sum([cross_entropy_loss(x, y) for (x, y) in zip(xL, yL)])
focal_list = lambda xL, yL: torch.sum([focal_loss(x, y, bce=True) for (x, y) in zip(xL, yL)])
classification_dice_list = lambda xL, yL: torch.sum(
    [classification_dice_loss(x, y, bce=True, background_weight=1) for (x, y) in zip(xL, yL)])


def cross_entropy_loss(gt, pred, weight=0.3, bce=False, background_weight=0):
    """loss_functions.py:26-44.  ``bce=True``: BCEWithLogits(input=pred, target=gt) (``weight`` and
    ``background_weight`` ignored, as there).  ``bce=False``: soft-label CE over dim 1 plus the
    background-weighted mirrored term."""
    if bce:
        return binary_cross_entropy(pred, gt)
    ops.nat.require_cuda(gt, pred)
    if gt.shape != pred.shape or pred.dim() < 2:
        raise ValueError("cross_entropy_loss(bce=False) expects gt and pred of the same [N,C,...] shape")
    n, c = pred.shape[0], pred.shape[1]
    a4, b4 = gt.reshape(n, c, 1, -1), pred.reshape(n, c, 1, -1)
    return ops.SoftCE.apply(a4, b4, float(background_weight))


def focal_loss(gt, pred, gamma=1.5, factor=0.1, background_weight=0):
    """loss_functions.py:46-50 (``gt`` is unused there too)."""
    return ops.leaf7(gt, pred, background_weight, scale=factor, shape=_shape(focal_gamma=gamma))[2]


def dice_loss(gt, pred, generalized=False, background_weight=1):
    """loss_functions.py:52-80."""
    out = ops.leaf7(gt, pred, background_weight, scale=1.0 / _M)
    return out[4] if generalized else out[3]


def twersky_loss(gt, pred, alpha=0.5, beta=0.3, background_weight=0):
    """loss_functions.py:82-94."""
    return ops.leaf7(gt, pred, background_weight, scale=1.0 / _M, shape=_shape(alpha=alpha, beta=beta))[5]


def focal_dice_coefficient(gt, pred, alpha=0.5, beta=0.3, gamma=1.8, background_weight=0):
    """loss_functions.py:96-108 (alpha/beta unused there too)."""
    return ops.leaf7(gt, pred, background_weight, scale=1.0 / _M, shape=_shape(focal_dice_gamma=gamma))[6]


def classification_dice_loss(gt, pred, factor=1e3, background_weight=1):
    """loss_functions.py:110-117 -> (dice, generalized dice, twersky, focal dice), each * factor * 0.33,
    from ONE statistics pass instead of four recomputations of the same sums."""
    out = ops.leaf7(gt, pred, background_weight, scale=(factor * 0.33) / _M)
    return out[3], out[4], out[5], out[6]
