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


class MulticlassLossStep:
    """The loss ``train()`` actually trains with, as ONE launch: ``train_multiclass.losses_fn(outputs, labels, ...)``
    for 3 organs is the plain sum over channels of the 7-loss leaf (train_multiclass.py:253-274; the composite branch is
    dead there), applied to ``F.sigmoid(net(x))`` (:134) and followed by ``loss.backward()`` (:147).

    step(logits, labels) -> (the 7 loss values, d(sum_k w_k loss_k)/d logits).  ``doubling`` = 1.0 for the
    train_multiclass flavour, 2.0 for loss_composite.losses_fn(composite_set_theory=False) (loss_composite.py:40).
    fp32 logits [N,3,H,W] with H*W % 4 == 0."""

    def __init__(self, weights, doubling=1.0, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.doubling = float(doubling)
        self.upstream = torch.tensor([float(w) for w in weights], dtype=torch.float32, device=self.device)

    def __call__(self, logits, labels, out=None):
        return ops.multiclass3_fused(logits, labels, self.doubling, self.upstream, out=out)


class LeafLossStep:
    """The plain (non-composite) ``losses_fn`` on LOGITS as ONE launch, for any organ count up to 64 and fp32 / bf16
    inputs -- in particular the single-organ configuration ORGANS=whole_body (BASELINE.json configs[0]):

    * C == 1 (loss_composite.py:32-40 / train_multiclass.py:264-274): the leaf with the PREDICTION in the gt slot,
      ``a = sigmoid(z)``, ``b = labels``, ``background_weight`` honoured;
    * C > 1 (loss_composite.py:28 / train_multiclass.py:260-262): sum over channels of the leaf
      ``a = labels_c, b = sigmoid(z_c)``; ``background_weight`` is dropped there, and here.

    step(logits, labels) -> (the 7 loss values, d(sum_k w_k loss_k)/d logits).  ``doubling`` = 2.0 for
    loss_composite.losses_fn (:40), 1.0 for the train_multiclass flavour."""

    def __init__(self, weights, doubling=2.0, background_weight=0.0, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.doubling = float(doubling)
        self.background_weight = float(background_weight)
        self.upstream = torch.tensor([float(w) for w in weights], dtype=torch.float32, device=self.device)

    def __call__(self, logits, labels, out=None):
        if logits.shape[1] == 1:
            losses, ga, _, _ = ops.pair_fused(logits, labels, ops.nat.FLAG_A_LOGIT, self.background_weight, self.doubling,
                                              self.upstream, want_a=True, want_b=False, out_a=out)
            return losses, ga
        losses, _, gb, _ = ops.pair_fused(labels, logits, ops.nat.FLAG_B_LOGIT, 0.0, self.doubling, self.upstream,
                                          want_a=False, want_b=True, out_b=out)
        return losses, gb


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


class PeerShardedCompositeLossStep(CompositeLossStep):
    """The data-parallel step as ONE cooperative launch per rank: statistics of the local shard -> in-kernel
    all-reduce of the 100 sums over NVLink peer memory (P2P stores into every peer's exchange buffer, release /
    acquire flags, fixed-order sum) -> closed forms -> gradient of the local shard.  No NCCL call on the step.

    torch.distributed is used once, at construction, to swap the 64-byte CUDA IPC handles of the exchange buffers.
    Every rank must call the step the same number of times (it is a collective).  Needs 16-byte aligned planes
    with H*W % 4 == 0 (otherwise use ShardedCompositeLossStep)."""

    def __init__(self, weights, group="world", **kw):
        import ctypes as C
        import torch.distributed as dist
        from . import distributed as dist_
        super().__init__(weights, **kw)
        self.pg = dist_._resolve(group)
        self.world = dist.get_world_size(self.pg)
        self.rank = dist.get_rank(self.pg)
        self.epoch = 0
        L = ops.nat.lib()
        dev = self.device.index
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        ops.nat.check(L.eco_xch_alloc(self.world, C.byref(own), handle, dev), "eco_xch_alloc")
        self._own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=self.pg)
        self._peers = []
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(self._own)
                continue
            p = C.c_void_p()
            ops.nat.check(L.eco_xch_open(h, C.byref(p), dev), f"eco_xch_open(rank {r})")
            self._peers.append(p.value)
            ptrs.append(p.value)
        self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        dist.barrier(group=self.pg)

    def __call__(self, logits, labels, out=None):
        import ctypes as C
        x, g = logits, labels
        ops.nat.require_cuda(x, g)
        if g.dtype != torch.float32:
            g = g.float()
        x, x_sn, x_sc = ops.nat.planes(x)
        g, g_sn, g_sc = ops.nat.planes(g)
        n, c, h, w = x.shape
        L = ops.nat.lib()
        ws = ops.nat.workspace("comp3", L.eco_composite3_ws_bytes(), x.device)
        losses = torch.empty((ops.nat.NLOSS,), dtype=torch.float32, device=x.device)
        gx = out if out is not None else torch.empty((n, c, h, w), dtype=x.dtype, device=x.device)
        vx, vg = ops.nat.view_of(x, x_sn, x_sc), ops.nat.view_of(g, g_sn, g_sc)
        og = ops.nat.out_of(gx, c * h * w, h * w)
        self.epoch += 1
        rc = L.eco_composite3_fused_sharded(C.byref(vx), C.byref(vg), n, h * w, int(self.from_logits),
                                            self.scales.data_ptr(), self.upstream.data_ptr(), ws.data_ptr(), ws.numel(),
                                            losses.data_ptr(), C.byref(og), self.peer_ptrs.data_ptr(), self.rank,
                                            self.world, self.epoch & 0xFFFFFFFF or 1, x.device.index,
                                            ops.nat.current_stream_ptr(x.device))
        ops.nat.check(rc, "eco_composite3_fused_sharded")
        return losses, gx

    def close(self):
        """Unmap the peers' buffers and free the own one (collective: all ranks should call it)."""
        import torch.distributed as dist
        if self._own is None:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.pg)
        L = ops.nat.lib()
        for p in self._peers:
            L.eco_xch_close(p, self.device.index)
        dist.barrier(group=self.pg)
        L.eco_xch_free(self._own, self.device.index)
        self._own, self._peers = None, []
