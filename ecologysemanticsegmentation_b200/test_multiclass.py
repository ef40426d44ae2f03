"""Drop-in for the scoring of ``ecology_semantic_segmentation/test_multiclass.py``.

The reference's ``test()`` (test_multiclass.py:30-108) does, per batch: ``out = F.sigmoid(net(x))`` (:58),
optionally the threshold rule ``out[out > T] = 1; out[out != 1] = 0`` (:68-69, commented there; live in
test_multiclass_sequential_densenetloss.py:88-89), then per class
``dice_loss(out_c, lab_c, background_weight=0)`` (:80-81) accumulated as a running sum and divided by the
number of batches (:82, :104) -- a mean of per-batch Dice, not a dataset-global Dice.

Here the sigmoid, the threshold, the integer pixel counts and the soft sums come from ONE kernel pass over
logits and labels (8 B/element), the counts are exact int64, and nothing syncs with the host until the end.
"""
from __future__ import annotations

import os

import torch

from . import ops
from . import distributed as dist_


def get_env_variable(name, default_value):
    """dataset/fish/__init__.py:10-14."""
    return os.environ.get(name, default_value)


ORGANS = [o for o in get_env_variable("ORGANS", "whole_body").split(",")]


def score_batch(logits, labels, threshold=None, *, group=None, inputs_are_probs=False, return_counts=False, ununion=False):
    """Per-class Dice of one batch, float32 [C] on the device.

    threshold=None: the live path of the reference -- soft Dice ``(2 sum p*lab + eps) / (sum (p + lab^2) + eps)``.
    threshold=T (float) or a sequence of up to 20 Ts: ``out = sigmoid(z) > float32(T)`` (strict), Dice from exact
    counts; a sequence returns [T, C].  ``group``: batch sharded across a process group, counts all-reduced.
    """
    ops.nat.require_cuda(logits, labels)
    thr_t = None
    many = False
    if threshold is not None:
        if isinstance(threshold, torch.Tensor):
            thr_t = threshold.to(device=logits.device, dtype=torch.float32).reshape(-1)
            many = threshold.dim() > 0
        else:
            many = hasattr(threshold, "__len__")
            vals = list(threshold) if many else [threshold]
            thr_t = torch.tensor([float(v) for v in vals], dtype=torch.float32, device=logits.device)
    if ununion and thr_t is not None:
        raise ValueError("ununion=True is the soft-Dice path of the sequential test; un-union in place "
                         "(subsets_union.return_union_sets_descending_order) before a thresholded score")
    counts, soft, inter = ops.dice_counts_ex(logits, labels, thr_t, inputs_are_probs=inputs_are_probs, ununion_preds=ununion)
    counts = dist_.allreduce_sums_(counts, group)
    soft = dist_.allreduce_sums_(soft, group)
    if inter is not None:
        inter = dist_.allreduce_sums_(inter, group)
    nthr = 0 if thr_t is None else thr_t.numel()
    dice, sdice = ops.dice_finalize(counts, soft, nthr, inter)
    out = sdice if thr_t is None else (dice if many else dice[0])
    if return_counts:
        return out, counts, soft
    return out


def to_uint8_masks(x, threshold=None, *, inputs_are_probs=False):
    """The byte tensors the reference dumps after scoring -- ``(t.numpy() * 255).astype(np.uint8)`` for the outputs
    ``t = sigmoid(net(x))`` (optionally thresholded, :68-69), the labels and the images (test_multiclass.py:90-92;
    test_video.py:129-130) -- produced on the device in one pass, so only 1 B/element crosses to the host.
    ``x``: logits, or probabilities / labels / images with ``inputs_are_probs=True``.  Returns uint8 CUDA [N,C,H,W]."""
    return ops.masks_u8(x, threshold, inputs_are_probs)


class DiceAccumulator:
    """Running ``test_dice = [[sum of per-batch Dice per class], number of batches]`` of test_multiclass.py:32,82,104,
    kept on the device."""

    def __init__(self):
        self.total = None
        self.count = 0

    def update(self, dice):
        self.total = dice.clone() if self.total is None else self.total + dice
        self.count += 1

    def mean(self):
        return self.total / float(self.count)


class StreamScorer:
    """Scores a stream of batches (the frame stream of test_video.py / the batch loop of test_multiclass.py:50-104)
    with ONE kernel launch per batch and nothing else: each batch's exact counts and soft sums land in their own
    slot of a device buffer, and -- because the per-batch Dice is only needed at the end (:104 takes the mean) --
    a sharded stream all-reduces the whole ``[batches, C, 3]`` buffer ONCE at the end instead of once per batch.
    ``result()`` = mean over batches of the per-batch per-class Dice, float32 [C] on the device, identical to
    ``score_stream`` (same counts, same float64 closed form)."""

    def __init__(self, n_classes, capacity, threshold=None, *, device=None, group=None):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.c, self.capacity, self.group = int(n_classes), int(capacity), group
        self.thr = None if threshold is None else torch.tensor([float(threshold)], dtype=torch.float32, device=device)
        self.counts = torch.zeros((self.capacity, self.c, 3), dtype=torch.int64, device=device)
        self.soft = torch.zeros((self.capacity, self.c, 3), dtype=torch.float64, device=device)
        self.inter = torch.zeros((self.capacity, self.c), dtype=torch.float64, device=device)   # sum out*lab, real label values
        self.n = 0

    def reset(self):
        self.n = 0

    def add(self, logits, labels, *, inputs_are_probs=False):
        if self.n >= self.capacity:
            raise IndexError(f"StreamScorer holds {self.capacity} batches")
        if logits.shape[1] != self.c:
            raise ValueError(f"expected {self.c} classes, got {logits.shape[1]}")
        ops.dice_counts_ex(logits, labels, self.thr, inputs_are_probs=inputs_are_probs,
                           out_counts=self.counts[self.n], out_soft=self.soft[self.n],
                           out_inter=self.inter[self.n] if self.thr is not None else None)
        self.n += 1

    def per_batch(self):
        """float32 [batches, C]: the Dice of every batch scored so far (all-reduced over ``group`` first)."""
        if self.n == 0:
            raise ValueError("no batch scored yet")
        counts, soft, inter = self.counts[:self.n], self.soft[:self.n], self.inter[:self.n]
        if dist_.world_size(self.group) > 1:
            counts = dist_.allreduce_sums_(counts.clone(), self.group)
            soft = dist_.allreduce_sums_(soft.clone(), self.group)
            inter = dist_.allreduce_sums_(inter.clone(), self.group)
        if self.thr is not None:
            # (2 sum out*lab + eps) / (sum out + sum lab^2 + eps) per batch and class, float64 like the closed form of
            # eco_dice_finalize_ex: a handful of [batches, C] element-wise ops at the end of the stream
            eps = 1e-7
            return ((2.0 * inter + eps) / (counts[..., 1].double() + soft[..., 2] + eps)).float()
        _, sdice = ops.dice_finalize(counts, soft.reshape(self.n * self.c, 3), 0)   # ... or on the class axis
        return sdice.reshape(self.n, self.c)

    def result(self):
        return self.per_batch().sum(0) / float(self.n)


def score_stream(batches, threshold=None, *, group=None):
    """Mean over batches of the per-batch per-class Dice (test_multiclass.py:104).  ``batches`` yields
    (logits, labels) CUDA tensor pairs."""
    acc = DiceAccumulator()
    for logits, labels in batches:
        acc.update(score_batch(logits, labels, threshold, group=group))
    return acc.mean()


def test(net, dataloader, models_dir="models/vgg", results_dir="test_results/", batch_size=1, saved_epoch=-1,
         single_model=False, threshold=None, ununion=False):
    """Same call signature and return value as the reference ``test()`` (test_multiclass.py:30): a float32
    CPU tensor [C] with the mean per-batch Dice per organ, or None when the results directory of this epoch
    already exists.  ``net`` maps images to logits; the sigmoid is fused into the scoring kernel.  Image dumping
    (``single_model``) is outside this path and not performed."""
    label_dirs = ORGANS
    dir_name = os.path.join(results_dir, "%s" % str(saved_epoch).zfill(4), ",".join(label_dirs))
    try:
        os.makedirs(dir_name)
    except Exception:
        if os.path.isdir(dir_name):
            print("Skipping epoch %d! Test already done!" % saved_epoch)
            return None
    net = net.eval()
    acc = DiceAccumulator()
    with torch.no_grad():
        for j, batch in enumerate(dataloader, 0):
            test_images, test_labels, image_ids = batch
            test_images = test_images.cuda()
            test_labels = test_labels.cuda()
            acc.update(score_batch(net(test_images), test_labels, threshold, ununion=ununion))
    dice_loss_val = acc.mean().cpu()
    print("Epoch %d: \n\t Test Dice Score: " % saved_epoch, dice_loss_val)
    print('Finished Testing')
    return dice_loss_val
