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

    step(logits, labels) -> (the 7 loss values f32 [7], d(sum_k w_k loss_k)/d logits)

    labels: float32, or uint8 / bool masks (1 B/element on the bus and in HBM; fp32 logits with H*W % 16 == 0).
    union_labels=True: ``labels`` are the raw per-organ masks and the label union train() applies before the loss
    (train_multiclass.py:110 -> utils/subsets_union.py:8-32, exclude_indices=[0]) is folded into the kernel's load stage.

    Pair weights and the numpy RNG (loss_composite.py:49-52): the reference draws 6 numbers per organ pair on EVERY call,
    also when ``early_stopped`` is False and the draws do not change the weights.  Here the weights are drawn at
    construction and, when ``early_stopped`` is True, again on every call (they change).  With ``early_stopped=False``
    the per-call draws are skipped unless ``advance_rng=True`` asks for the reference's exact consumption of the global
    numpy stream (18 draws per call, ~0.1 ms of host time)."""

    def __init__(self, weights, relative_set_ratios=DEFAULT_RATIOS, early_stopped=False, from_logits=True,
                 device=None, union_labels=False, advance_rng=False):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ratios = list(relative_set_ratios)
        self.early_stopped = early_stopped
        self.from_logits = from_logits
        self.union_labels = bool(union_labels)
        self.advance_rng = bool(advance_rng)
        self.upstream = torch.tensor([float(w) for w in weights], dtype=torch.float32, device=self.device)
        self.scales = torch.empty(ops.nat.C3_NLEAF, dtype=torch.float64, device=self.device)
        # two pinned staging buffers, alternated, each guarded by the event of its last upload: the host never rewrites
        # a buffer whose asynchronous copy may not have run yet
        self._host_scales = [torch.empty(ops.nat.C3_NLEAF, dtype=torch.float64).pin_memory() for _ in range(2)]
        self._host_events = [None, None]
        self._host_turn = 0
        self._first = True
        self.redraw()

    def redraw(self):
        """Draw the pair weights with the reference's numpy RNG stream (loss_composite.py:49-52) and upload them."""
        sc = composite3_leaf_scales(draw_pair_weights(self.ratios, self.early_stopped))
        k = self._host_turn
        self._host_turn ^= 1
        if self._host_events[k] is not None:
            self._host_events[k].synchronize()
        self._host_scales[k].copy_(torch.tensor(sc, dtype=torch.float64))
        with torch.cuda.device(self.device):
            self.scales.copy_(self._host_scales[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self._host_events[k] = ev

    def _per_call_draws(self):
        if self._first:          # the construction-time draw serves the first call
            self._first = False
        elif self.early_stopped:
            self.redraw()
        elif self.advance_rng:
            draw_pair_weights(self.ratios, False)   # same 18 draws as the reference; the weights do not change

    def _prepared(self, logits, labels, peers=None):
        """Cached launch block for this pair of buffers (None when the general path has to serve the call)."""
        cache = self.__dict__.setdefault("_launches", {})
        key = (logits.data_ptr(), labels.data_ptr(), torch._C._cuda_getCurrentRawStream(self.device.index))
        ent = cache.get(key)
        if ent is not None and ent.sig == ops.PreparedComposite3.signature(logits, labels):
            return ent
        if ent is False:
            return None
        try:
            ent = ops.PreparedComposite3(logits, labels, self.scales, self.upstream, self.from_logits, self.union_labels, peers)
            if ent.L.eco_composite3_ws_bytes() > ent.ws.numel():
                raise ValueError("workspace")
        except ValueError:
            ent = False
        if len(cache) >= 64:
            cache.clear()
        cache[key] = ent
        return ent or None

    def __call__(self, logits, labels, out=None):
        self._per_call_draws()
        if torch.cuda.current_device() == self.device.index and not self.union_labels_needs_fallback(logits):
            ent = self._prepared(logits, labels)
            if ent is not None:
                return ent.run(out)
        losses, grad = ops.composite3_fused(logits, labels, self.scales, self.upstream, self.from_logits, out=out,
                                            union_labels=self.union_labels)
        return losses, grad

    def union_labels_needs_fallback(self, logits):
        """The prepared launch serves logits (fp32 / bf16); probabilities take the general path."""
        return logits.dtype not in (torch.float32, torch.bfloat16) or not self.from_logits

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
        self._per_call_draws()
        if self.union_labels:   # the two-kernel path has no fused union: the in-place kernel on a float copy
            from .subsets_union import return_union_sets_descending_order
            labels = return_union_sets_descending_order(labels.float().clone())
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
    Every rank must call the step the same number of times (it is a collective, and like one it waits for the slowest
    rank: while it waits the cooperative grid holds all SMs of the device).  A peer that does not show up within
    ``timeout_ms`` poisons the step's outputs with NaN and sets a status word that the NEXT call (or ``check()``) turns
    into an EcoLossError.  Needs 16-byte aligned planes with H*W % 4 == 0 (otherwise use ShardedCompositeLossStep)."""

    def __init__(self, weights, group="world", timeout_ms=30000.0, **kw):
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
        # time-out word in mapped pinned host memory: the kernel sets it, the host reads it without synchronising
        self._status = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.timeout_ms = float(timeout_ms)
        dist.barrier(group=self.pg)

    def check(self):
        """Raise if a wait on a peer timed out in an earlier step (its outputs were poisoned with NaN)."""
        if int(self._status[0]) != 0:
            self._status[0] = 0
            raise ops.nat.EcoLossError(
                f"rank {self.rank}: a peer did not deliver its partial sums within {self.timeout_ms / 1e3:.1f} s "
                "(in-kernel NVLink exchange); every rank must call the step the same number of times")

    def __call__(self, logits, labels, out=None):
        self.check()
        self._per_call_draws()
        self.epoch += 1
        peers = self.__dict__.get("_peers_struct")
        if peers is None:
            peers = self._peers_struct = ops.nat.EcoPeerExchange(self.peer_ptrs.data_ptr(), self.rank, self.world, 1, 0,
                                                                 self._status.data_ptr(), self.timeout_ms)
        peers.epoch = self.epoch & 0xFFFFFFFF or 1
        if torch.cuda.current_device() == self.device.index and not self.union_labels_needs_fallback(logits):
            ent = self._prepared(logits, labels, peers)
            if ent is not None:
                return ent.run(out)
        return ops.composite3_fused(logits, labels, self.scales, self.upstream, self.from_logits, out=out,
                                    union_labels=self.union_labels, peers=peers)

    def close(self):
        """Unmap the peers' buffers and free the own one (collective: all ranks should call it)."""
        import torch.distributed as dist
        if self._own is None:
            return
        torch.cuda.synchronize(self.device)
        self.check()
        dist.barrier(group=self.pg)
        L = ops.nat.lib()
        for p in self._peers:
            L.eco_xch_close(p, self.device.index)
        dist.barrier(group=self.pg)
        L.eco_xch_free(self._own, self.device.index)
        self._own, self._peers = None, []
