"""Single-launch training-step API: loss values AND the gradient of the weighted loss in one pass.

``train()`` in the reference computes ``outputs = F.sigmoid(net(inputs))`` (train_multiclass.py:133-134),
calls ``losses_fn`` (:139-141), forms ``loss = focal_dice_w*focal_dice + bce_l_w*bce_l + generalized_dice_w*
(generalized_dice + twersky_dice)`` (:145) and calls ``loss.backward()`` (:147).  When the 0/1 weights of :145
are known before the call -- they are: they depend on the epoch only (:92-100) -- the whole chain
sigmoid -> 21-leaf composite loss -> d loss / d logits fits one cooperative kernel launch.
"""
from __future__ import annotations

import torch

from . import ops
from .loss_composite import DEFAULT_RATIOS, LossList, composite3_leaf_scales, draw_pair_weights

LOSS_NAMES = ("ce", "bce", "focal", "dice", "generalized_dice", "twersky", "focal_dice")


def loss_weights(ce=0.0, bce=0.0, focal=0.0, dice=0.0, generalized_dice=0.0, twersky=0.0, focal_dice=0.0):
    """Upstream weights in the loss order of loss_composite.py:39."""
    return [ce, bce, focal, dice, generalized_dice, twersky, focal_dice]


class CompositeLossStep:
    """Callable that owns the small device-side parameter buffers so a step is launch-only (CUDA-graph friendly).

    step(logits, labels) -> (LossList of the 7 loss values, d(sum_k w_k loss_k)/d logits)
    """

    def __init__(self, weights, relative_set_ratios=DEFAULT_RATIOS, early_stopped=False, from_logits=True,
                 device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ratios = list(relative_set_ratios)
        self.early_stopped = early_stopped
        self.from_logits = from_logits
        self.upstream = torch.tensor([float(w) for w in weights], dtype=torch.float32, device=self.device)
        self.scales = torch.empty(ops.nat.C3_NLEAF, dtype=torch.float64, device=self.device)
        self._host_scales = torch.empty(ops.nat.C3_NLEAF, dtype=torch.float64).pin_memory()
        self.redraw()

    def redraw(self):
        """Draw the pair weights with the reference's numpy RNG stream (loss_composite.py:49-52) and upload them."""
        sc = composite3_leaf_scales(draw_pair_weights(self.ratios, self.early_stopped))
        self._host_scales.copy_(torch.tensor(sc, dtype=torch.float64))
        self.scales.copy_(self._host_scales, non_blocking=True)

    def __call__(self, logits, labels, out=None):
        losses, grad = ops.composite3_fused(logits, labels, self.scales, self.upstream, self.from_logits, out=out)
        return losses, grad

    def as_losslist(self, losses):
        return LossList(losses.unbind(0))


class ShardedCompositeLossStep(CompositeLossStep):
    """Batch sharded over a process group (one process per GPU): statistics kernel -> ONE all-reduce of the
    100 float64 sums (800 B) over NCCL/NVLink -> closed forms -> gradient kernel for this rank's shard.
    The result equals the single-device loss of the concatenated batch."""

    def __init__(self, weights, group="world", **kw):
        super().__init__(weights, **kw)
        self.group = group

    def __call__(self, logits, labels, out=None):
        from . import distributed as dist_
        acc = ops.composite3_stats(logits, labels, self.from_logits)
        acc = dist_.allreduce_sums_(acc, self.group)
        losses, jac, _ = ops.composite3_finalize(acc, self.scales)
        grad = ops.composite3_grad(logits, labels, self.from_logits, jac, self.upstream, out=out)
        return losses, grad
