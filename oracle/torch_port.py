"""Op-order-faithful CPU restatement of the reference loss path in torch eager ops.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the reference
lines it follows; ``ess`` abbreviates ``ecology_semantic_segmentation``.  The op order
inside each expression is the reference's, so on the same device and dtype the results
(and the autograd gradients) are bit-identical to the unmodified reference --
``tests/test_oracle_vs_reference.py`` asserts exactly that.

Everything is written on *slots*: ``a`` is the reference's first positional argument
("gt" slot) and ``b`` the second ("pred" slot).  Which of the two is really the label
depends on the caller.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-7
DEFAULT_RATIOS = (1.0, 0.43197708, 0.22319692)  # ess/loss_composite.py:21


# ----------------------------------------------------------------------------------------------
# primitives (ess/loss_functions.py)
# ----------------------------------------------------------------------------------------------
def pair_bce(a, b):
    """ess/loss_functions.py:34,44 with ess/__init__.py:24.

    ``BCEWithLogitsLoss()(input=b, target=a)``: the *second* slot is fed as the logit even
    though callers pass probabilities or labels there (Appendix A item 1).  The outer
    ``torch.mean`` of a 0-d tensor is kept because the reference has it.
    """
    return torch.mean(F.binary_cross_entropy_with_logits(b, a))


def pair_soft_ce(a, b, background_weight=0):
    """ess/loss_functions.py:29-30,44 -- ``F.cross_entropy`` with probability targets over dim 1."""
    ce = F.cross_entropy(b, a)
    ce = ce + background_weight * F.cross_entropy(1 - b, 1 - a)
    return torch.mean(ce)


def pair_focal(a, b, gamma=1.5, factor=0.1, background_weight=0):
    """ess/loss_functions.py:46-50.  ``a`` is unused by the reference as well."""
    fg = -torch.pow((1 - b), gamma) * torch.log(b + EPS)
    fg = fg + (-background_weight * torch.pow(b, gamma) * torch.log(1 - b + EPS))
    return factor * torch.mean(fg)


def pair_dice(a, b, generalized=False, background_weight=1):
    """ess/loss_functions.py:52-80 (both branches)."""
    if not generalized:
        num = 2 * torch.sum(a * b)
        den = torch.sum(a + b * b)
        fg = (num + EPS) / (den + EPS)
        num_bg = 2 * torch.sum((1 - a) * (1 - b))
        den_bg = 2 * torch.sum((1 - a) + (1 - b) * (1 - b))  # the extra factor 2 is the reference's (:60)
        bg = (num_bg + EPS) / (den_bg + EPS)
        return -fg - background_weight * bg
    a0, b0 = (1 - a), (1 - b)
    num = torch.sum(a * b) + EPS
    den = torch.sum(a + b * b) + EPS
    dc = num / den
    num_bg = torch.sum(a0 * b0) + EPS
    den_bg = torch.sum(a0 + b0 * b0) + EPS
    dc = dc + background_weight * (num_bg / den_bg)
    return -dc


def pair_tversky(a, b, alpha=0.5, beta=0.3, background_weight=0):
    """ess/loss_functions.py:82-94 ("twersky_loss")."""
    tp = torch.sum(a * b)
    den = torch.sum(a * b) + alpha * torch.sum((1 - b) * a) + beta * torch.sum(b * (1 - a))
    fg = -(tp + EPS) / (den + EPS)
    a = 1 - a
    b = 1 - b
    tp_bg = torch.sum(a * b)
    den_bg = torch.sum(a * b) + alpha * torch.sum((1 - b) * a) + beta * torch.sum(b * (1 - a))
    bg = -(tp_bg + EPS) / (den_bg + EPS)
    return fg + background_weight * bg


def pair_focal_dice(a, b, gamma=1.8, background_weight=0):
    """ess/loss_functions.py:96-108 (alpha/beta are unused there too)."""
    num = 2 * torch.sum(a * b)
    den = torch.sum(a + b * b)
    dc = (num + EPS) / (den + EPS)
    fg = -torch.pow(1 - dc, gamma) * torch.log(dc + EPS)
    num_bg = 2 * torch.sum((1 - a) * (1 - b))
    den_bg = torch.sum((1 - a) + (1 - b) * (1 - b))
    dc_bg = (num_bg + EPS) / (den_bg + EPS)
    bg = -torch.pow(1 - dc_bg, gamma) * torch.log(dc_bg + EPS)
    return fg + background_weight * bg


def pair_dice_family(a, b, factor=1e3, background_weight=1):
    """ess/loss_functions.py:110-117 -> (dice, generalized dice, tversky, focal dice) * factor * 0.33."""
    d = pair_dice(a, b, background_weight=background_weight)
    gd = pair_dice(a, b, generalized=True, background_weight=background_weight)
    tv = pair_tversky(a, b, background_weight=background_weight)
    fd = pair_focal_dice(a, b, background_weight=background_weight)
    m = factor * 0.33
    return d * m, gd * m, tv * m, fd * m


# ----------------------------------------------------------------------------------------------
# the 7-loss leaf and the two ``losses_fn`` flavours
# ----------------------------------------------------------------------------------------------
def leaf7(a, b, background_weight=0, double_up=True):
    """The single-channel branch: ess/loss_composite.py:32-40 (``double_up=True``, the ``+=`` at :40)
    or ess/train_multiclass.py:269-274 (``double_up=False``).  Order of the 7:
    [ce, bce, focal, dice, generalized dice, tversky, focal dice]."""
    bce = pair_bce(a, b)
    ce = pair_soft_ce(a, b, background_weight=background_weight)
    fl = pair_focal(a, b, factor=1, background_weight=background_weight)
    d, gd, tv, fd = pair_dice_family(a, b, factor=10, background_weight=background_weight)
    out = [ce, bce, fl, d, gd, tv, fd]
    if double_up:
        out = [v + v for v in out]
    return out


def union_operand(sp, p):
    """ess/loss_composite.py:94 -- u(sp, p), evaluated in the reference's op order."""
    return sp * (1 - p) + (sp * p + p) * 0.5


def _per_channel(x, g, double_up):
    """C>1 recursion with swapped slots: ess/loss_composite.py:28-30 / ess/train_multiclass.py:260-262.
    ``background_weight`` is silently dropped there (Appendix A item 4)."""
    per_c = [leaf7(g[:, c:c + 1, :, :], x[:, c:c + 1, :, :], 0, double_up) for c in range(g.shape[1])]
    return [sum(col) for col in zip(*per_c)]


def composite_pair_terms(x, g, i, j):
    """The six (a, b) operand pairs of ess/loss_composite.py:56-81 for organ pair i<j, in order,
    tagged with which weight they take ("j", "i", "d", "i", "d", "iij")."""
    # Every mention re-slices, as the reference does, so that the autograd graph (and hence the
    # fp32 accumulation order of the gradient) is the same as the reference's.
    X = lambda c: x[:, c:c + 1, ...]
    G = lambda c: g[:, c:c + 1, ...]
    return [
        ("j", X(i) * X(j), G(j)),                                                     # :56  intersection
        ("i", G(i), union_operand(X(i), X(j))),                                       # :59  union
        ("d", X(i) * torch.abs(X(i) - X(j)), torch.abs(G(i) - G(j))),                 # :63-65
        ("i", G(i), union_operand(X(i), torch.abs(X(i) - X(j)))),                     # :68-70
        ("d", X(i) * (torch.abs(X(i) - X(j)) * X(i)), torch.abs(G(i) - G(j))),        # :74-76
        ("iij", G(i), union_operand(X(i), torch.abs(X(i) - X(j)) * X(i))),            # :79-81
    ]


def draw_pair_weights(ratios, i, j, early_stopped):
    """ess/loss_composite.py:49-52.  Six numpy global-RNG draws per pair, taken even when
    ``early_stopped`` is False (Appendix A item 8)."""
    es = int(early_stopped)
    w_i = (1 / ratios[i]) * (1 - es * np.random.choice([0, 1]) * np.random.rand())
    w_j = (1 / ratios[j]) * (1 - es * np.random.choice([0, 1]) * np.random.rand())
    w_d = (1 / (ratios[i] - ratios[j])) * (1 - es * np.random.choice([0, 1]) * np.random.rand())
    return w_i, w_j, w_d


def losses_composite(x, g, composite_set_theory=False, background_weight=0, early_stopped=False,
                     relative_set_ratios=DEFAULT_RATIOS):
    """``ess/loss_composite.py:21-84`` ``losses_fn``."""
    assert x.shape[1] == len(relative_set_ratios) or not composite_set_theory, "Organ ratios size mismatch!"
    if g.shape[1] > 1:
        total = _per_channel(x, g, True)
    else:
        total = leaf7(x, g, background_weight, True)
    if composite_set_theory:
        C = g.shape[1]
        for i in range(C - 1):
            for j in range(i + 1, C):
                w_i, w_j, w_d = draw_pair_weights(relative_set_ratios, i, j, early_stopped)
                for tag, a, b in composite_pair_terms(x, g, i, j):
                    term = leaf7(a, b, 0, True)
                    if tag == "j":
                        term = [v * w_j for v in term]
                    elif tag == "i":
                        term = [v * w_i for v in term]
                    elif tag == "d":
                        term = [v * w_d for v in term]
                    else:  # w_i * w_i * w_j applied as three successive list scalings (:81)
                        term = [((v * w_i) * w_i) * w_j for v in term]
                    total = [s + t for s, t in zip(total, term)]
    return total


def losses_train_multiclass(x, g, composite_set_theory=False, background_weight=0, early_stopped=False):
    """``ess/train_multiclass.py:253-303`` ``losses_fn``: no doubling; the composite branch is
    unreachable for C>1 (early return :262) and raises for C=1."""
    if g.shape[1] > 1:
        return _per_channel(x, g, False)
    out = leaf7(x, g, background_weight, False)
    if composite_set_theory:
        raise ValueError("train_multiclass.losses_fn composite branch is broken in the reference (C=1)")
    return out


# ----------------------------------------------------------------------------------------------
# evaluation (ess/test_multiclass.py:58-82,104)
# ----------------------------------------------------------------------------------------------
def threshold_inplace(out, threshold):
    """ess/test_multiclass.py:68-69 (commented there) / test_multiclass_sequential_densenetloss.py:88-89."""
    out[out > threshold] = 1
    out[out != 1] = 0
    return out


def eval_batch_dice(logits, labels, threshold=None):
    """One batch of ``test()``: sigmoid (:58), optional threshold rule, per-class
    ``-dice_loss(out_c, lab_c, background_weight=0)`` (:80-82).  Returns a python list of 0-d tensors."""
    out = torch.sigmoid(logits)
    if threshold is not None:
        out = threshold_inplace(out, threshold)
    return [-pair_dice(out[:, c:c + 1, :, :], labels[:, c:c + 1, :, :], background_weight=0)
            for c in range(labels.shape[1])]


def eval_stream_dice(batches, threshold=None):
    """Mean over batches of the per-batch Dice (:82,:104)."""
    acc, count = None, 0
    for logits, labels in batches:
        d = eval_batch_dice(logits, labels, threshold)
        acc = [0 - (-v) for v in d] if acc is None else [s - (-v) for s, v in zip(acc, d)]
        count += 1
    return torch.tensor([float(v) for v in acc]) / float(count)


# ----------------------------------------------------------------------------------------------
# steps adjacent to the path
# ----------------------------------------------------------------------------------------------
def union_sets_descending(ann, exclude_indices=(0,), reverse=False):
    """ess/utils/subsets_union.py:8-32 -- in place, along the CLASS dim (dim 1) of [N,C,H,W].
    forward: every non-excluded channel but the last becomes the sum of itself and all later channels, then
    everything above 1 is clamped to 1; reverse: from the back, channel c becomes |c - (c+1)|."""
    C = ann.shape[1]
    if not reverse:
        for c in range(C - 1):
            if c in exclude_indices:
                continue
            ann[:, c] = torch.sum(ann[:, c:], axis=1)
        ann[ann > 1] = 1
    else:
        for c in range(C - 2, -1, -1):
            if c in exclude_indices:
                continue
            ann[:, c] = torch.abs(ann[:, c] - ann[:, c + 1])
    return ann


def union_sets_descending_batchdim(ann, exclude_indices=(0,)):
    """ess/train_multiclass.py:32-45 -- the twin that train() really calls (:110): same recipe but along dim 0
    (the batch dimension of the [N,C,H,W] labels; a bug of the reference that a drop-in has to keep)."""
    for i in range(ann.shape[0] - 1):
        if i in exclude_indices:
            continue
        ann[i] = sum(x for x in ann[i:])
    ann[ann > 1] = 1
    return ann


def losses_sequential_densenet(x, g, composite_set_theory=False, background_weight=0, early_stopped=False):
    """ess/train_multiclass_sequential_densenetloss.py:272-362 ``losses_fn``: per-channel leaves without doubling,
    plus -- for C>1 -- the extra leaf(a = g_1 - g_2, b = |x_1 - x_2|) added to channel 1 (:285).  The composite
    branch is unreachable for C>1 and hits an undefined name for C=1."""
    if g.shape[1] > 1:
        per_c = [leaf7(g[:, c:c + 1, :, :], x[:, c:c + 1, :, :], 0, False) for c in range(g.shape[1])]
        extra = leaf7(g[:, 1:2, :, :] - g[:, 2:3, :, :], torch.abs(x[:, 1:2, :, :] - x[:, 2:3, :, :]), 0, False)
        per_c[1] = [a + b for a, b in zip(extra, per_c[1])]
        return [sum(col) for col in zip(*per_c)]
    out = leaf7(x, g, background_weight, False)
    if composite_set_theory:
        # :304-320 slice channels 1 and 2 of a 1-channel tensor; the first leaf on them raises inside BCEWithLogits
        raise ValueError("Target size must be the same as input size")
    return out
