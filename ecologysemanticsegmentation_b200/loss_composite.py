"""Drop-in for ``ecology_semantic_segmentation/loss_composite.py``: ``LossList``, ``losses_fn``,
``intersection_loss``, ``union_loss`` with the reference's signatures and quirks.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

DEFAULT_RATIOS = [1., 0.43197708, 0.22319692]


class LossList(list):
    """loss_composite.py:9-17: ``+=`` adds element-wise (same length required) and returns a new LossList;
    ``* w`` scales element-wise and insists on a python/numpy *float* weight."""

    def __iadd__(self, other_list):
        assert len(self) == len(other_list), \
            "Lists to be added need the same length! (%d vs %d)" % (len(self), len(other_list))
        return LossList([x + y for x, y in zip(self, other_list)])

    def __mul__(self, w):
        assert isinstance(w, float), "Multiplication supported for numerical weights only! Found %s" % type(w)
        return LossList([x * w for x in self])


def draw_pair_weights(relative_set_ratios, early_stopped):
    """The host-side weight draw of loss_composite.py:46-52, verbatim in its use of the GLOBAL numpy RNG:
    two draws (choice, rand) per weight, three weights per pair in the order w_idx, w_jdx, w_diff, taken even
    when ``early_stopped`` is False.  Returns [(i, j, w_i, w_j, w_d), ...]."""
    # ``np.random.choice([0, 1])`` is written as ``np.random.randint(0, 2)``: the legacy RandomState implements the former
    # as exactly that call after converting the list to an array, so value and stream consumption are identical (asserted
    # in tests/test_host_logic.py) at a third of the host time -- this runs on every losses_fn call.
    out = []
    coin, rand = np.random.randint, np.random.rand
    es = int(early_stopped)
    length = len(relative_set_ratios)
    for idx in range(length - 1):
        for jdx in range(idx + 1, length):
            w_idx = (1 / relative_set_ratios[idx]) * (1 - es * coin(0, 2) * rand())
            w_jdx = (1 / relative_set_ratios[jdx]) * (1 - es * coin(0, 2) * rand())
            w_diff = (1 / (relative_set_ratios[idx] - relative_set_ratios[jdx])) * (1 - es * coin(0, 2) * rand())
            out.append((idx, jdx, w_idx, w_jdx, w_diff))
    return out


def composite3_leaf_scales(pair_weights, doubling=2.0):
    """21 leaf scales in the kernel's leaf order: 3 channel leaves, then per pair I1,U1,I2,U2,I3,U3
    (weights w_j, w_i, w_d, w_i, w_d, w_i*w_i*w_j of loss_composite.py:56-81; ``doubling`` is the ``+=`` at :40)."""
    scales = [doubling] * 3
    for (_, _, w_i, w_j, w_d) in pair_weights:
        scales += [doubling * w_j, doubling * w_i, doubling * w_d, doubling * w_i, doubling * w_d,
                   doubling * w_i * w_i * w_j]
    return scales


def _per_channel(x, g, doubling, group=None, from_logits=False, key=None):
    """C>1 recursion (loss_composite.py:28-30): leaf(a = g_c, b = x_c) summed over channels;
    ``background_weight`` is dropped there."""
    flags = ops.nat.FLAG_B_LOGIT if from_logits else 0
    return LossList(ops.PairLeaves.apply(g, x, 0.0, float(doubling), flags, group, None, key))


def losses_fn(x, g, composite_set_theory=False, background_weight=0, early_stopped=False,
              relative_set_ratios=DEFAULT_RATIOS, *, group=None, from_logits=False, _site="lc"):
    """loss_composite.py:21-84.  Returns ``LossList[ce, bce, focal, dice, generalized_dice, twersky, focal_dice]``.

    Extensions (keyword-only, default to the reference behaviour): ``group`` shards the batch over a
    torch.distributed process group (sums all-reduced); ``from_logits`` fuses the sigmoid of
    train_multiclass.py:134 into the kernels (``x`` are logits, gradients are w.r.t. the logits)."""
    CLASS_INDEX = 1
    assert x.shape[CLASS_INDEX] == len(relative_set_ratios) or not composite_set_theory, "Organ ratios size mismatch!"
    ops.nat.require_cuda(x, g)
    C = g.shape[CLASS_INDEX]

    labels_need_grad = g.requires_grad and torch.is_grad_enabled()
    if composite_set_theory and C == 3 and x.dim() == 4 and not labels_need_grad:
        weights = draw_pair_weights(relative_set_ratios, early_stopped)
        scales = composite3_leaf_scales(weights)
        return LossList(ops.Composite3.apply(x, g, scales, bool(from_logits), group))

    if composite_set_theory and from_logits:
        # the pair-by-pair composition below works on probabilities (organ counts other than 3, or labels that
        # require grad): one explicit sigmoid, exactly the reference's F.sigmoid at train_multiclass.py:134
        x, from_logits = torch.sigmoid(x), False

    # A top-level call of the plain loss runs as one launch with the upstream weights of the previous step anticipated
    # (ops.PairLeaves, `key`); the pair-by-pair composition below makes 1 + 6 * pairs leaf calls whose upstream weights differ
    # from call to call (intersection_loss / union_loss pass _site=None), so it keeps the three pair-leaf launches per leaf.
    key = None if (composite_set_theory or _site is None) else _site
    if C > 1:
        return_losses = _per_channel(x, g, 2.0, group, from_logits, key)
    else:
        # single channel: prediction goes in the gt slot, background_weight is honoured, result doubled (:32-40)
        flags = ops.nat.FLAG_A_LOGIT if from_logits else 0
        return_losses = LossList(ops.leaf7(x, g, background_weight, scale=2.0, flags=flags, group=group,
                                           key=None if key is None else key + "1"))

    if composite_set_theory:
        # generic organ count (and labels that require grad): the reference's own composition, each leaf one
        # pass of the pair-leaf kernels, gradients w.r.t. BOTH slots
        for (idx, jdx, w_idx, w_jdx, w_diff) in draw_pair_weights(relative_set_ratios, early_stopped):
            xi, xj = x[:, idx:idx + 1, ...], x[:, jdx:jdx + 1, ...]
            gi, gj = g[:, idx:idx + 1, ...], g[:, jdx:jdx + 1, ...]
            return_losses += intersection_loss(xi, xj, gj, group=group) * w_jdx
            return_losses += union_loss(xi, xj, gi, group=group) * w_idx
            return_losses += intersection_loss(xi, torch.abs(xi - xj), torch.abs(gi - gj), group=group) * w_diff
            return_losses += union_loss(xi, torch.abs(xi - xj), gi, group=group) * w_idx
            return_losses += intersection_loss(xi, torch.abs(xi - xj) * xi, torch.abs(gi - gj), group=group) * w_diff
            return_losses += union_loss(xi, torch.abs(xi - xj) * xi, gi, group=group) * w_idx * w_idx * w_jdx
    return return_losses


def intersection_loss(superset_p, set_p, set_g, *, group=None):
    """loss_composite.py:87-88 -- ``losses_fn(superset_p * set_p, set_g)``."""
    return LossList(losses_fn(superset_p * set_p, set_g, composite_set_theory=False, group=group, _site=None))


def union_loss(superset_p, set_p, superset_g, *, group=None):
    """loss_composite.py:92-94 -- ``losses_fn(superset_g, u)`` with u evaluated in the reference's op order."""
    return LossList(losses_fn(superset_g, (superset_p * (1 - set_p) + (superset_p * set_p + set_p) * 0.5),
                              composite_set_theory=False, group=group, _site=None))
