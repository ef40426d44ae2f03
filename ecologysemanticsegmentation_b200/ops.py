"""Tensor-level wrappers over the C ABI and the autograd Functions built on them.

Loss order everywhere: [ce, bce, focal, dice, generalized_dice, twersky, focal_dice]
(ecology_semantic_segmentation/loss_composite.py:39).  "Slot a" is the reference's first positional
argument, "slot b" the second.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as nat
from . import distributed as dist_

M_DICE = 10 * 0.33  # classification_dice_loss(factor=10), loss_functions.py:116


# ------------------------------------------------------------------------------------------------
# raw launches
# ------------------------------------------------------------------------------------------------
def _dev(t):
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _shape_ref(shape):
    """None or (focal_gamma, tversky_alpha, tversky_beta, focal_dice_gamma) -> ctypes argument (NULL = defaults)."""
    if shape is None or tuple(shape) == nat.DEFAULT_SHAPE:
        return None
    return C.byref(nat.EcoLeafShape(*[float(v) for v in shape]))


def pair_stats(a, b, flags=0, shape=None):
    """a, b: [N,C,H,W] CUDA tensors -> float64 [C, 8] sums of the C (a_c, b_c) leaves of this shard."""
    nat.require_cuda(a, b)
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError(f"pair_stats expects two [N,C,H,W] tensors of equal shape, got {tuple(a.shape)} and {tuple(b.shape)}")
    a, a_sn, a_sc = nat.planes(a)
    b, b_sn, b_sc = nat.planes(b)
    n, c, h, w = a.shape
    L = nat.lib()
    ws = nat.workspace("pair", L.eco_pair_ws_bytes(c), a.device)
    sums = torch.empty((c, nat.NSTAT), dtype=torch.float64, device=a.device)
    va, vb = nat.view_of(a, a_sn, a_sc), nat.view_of(b, b_sn, b_sc)
    rc = L.eco_pair_stats_shaped(C.byref(va), C.byref(vb), n, c, h * w, flags, _shape_ref(shape), ws.data_ptr(),
                                 ws.numel(), sums.data_ptr(), _dev(a), nat.current_stream_ptr(a.device))
    nat.check(rc, "eco_pair_stats")
    return sums


def pair_finalize(sums, background_weight, scales, shape=None):
    """sums float64 [C,8] -> (losses f32 [C,7], total f32 [7], jac f64 [C,7,7])."""
    c = sums.shape[0]
    dev = sums.device
    losses = torch.empty((c, nat.NLOSS), dtype=torch.float32, device=dev)
    total = torch.empty((nat.NLOSS,), dtype=torch.float32, device=dev)
    jac = torch.empty((c, nat.NLOSS, nat.NJAC), dtype=torch.float64, device=dev)
    sc = (C.c_double * c)(*[float(s) for s in scales])
    rc = nat.lib().eco_pair_finalize_shaped(sums.data_ptr(), c, float(background_weight), sc, _shape_ref(shape),
                                            losses.data_ptr(), total.data_ptr(), jac.data_ptr(), _dev(sums),
                                            nat.current_stream_ptr(dev))
    nat.check(rc, "eco_pair_finalize")
    return losses, total, jac


def pair_grad(a, b, flags, jac, upstream, want_a, want_b, shape=None):
    a, a_sn, a_sc = nat.planes(a)
    b, b_sn, b_sc = nat.planes(b)
    n, c, h, w = a.shape
    ga = torch.empty((n, c, h, w), dtype=a.dtype, device=a.device) if want_a else None
    gb = torch.empty((n, c, h, w), dtype=b.dtype, device=b.device) if want_b else None
    va, vb = nat.view_of(a, a_sn, a_sc), nat.view_of(b, b_sn, b_sc)
    oa, ob = nat.out_of(ga, c * h * w, h * w), nat.out_of(gb, c * h * w, h * w)
    rc = nat.lib().eco_pair_grad_shaped(C.byref(va), C.byref(vb), n, c, h * w, flags, _shape_ref(shape), jac.data_ptr(),
                                        upstream.data_ptr(), C.byref(oa), C.byref(ob), 0, _dev(a),
                                        nat.current_stream_ptr(a.device))
    nat.check(rc, "eco_pair_grad")
    return ga, gb


def pair_fused(a, b, flags, background_weight, scale, upstream, want_a=False, want_b=True, out_a=None, out_b=None,
               shape=None, upstream_prev=None, losses=None, sums=None):
    """ONE cooperative launch for a step of C independent leaves (eco_pair_fused): sums -> closed forms -> gradient of
    sum_k upstream[k] * loss_k w.r.t. slot a and/or slot b (w.r.t. the logits where a slot is flagged as logits).
    Returns (losses f32 [7] summed over the channels, ga, gb, sums f64 [C, 8]).
    ``upstream_prev`` (with the ``losses`` / ``sums`` / ``out_*`` buffers of the step it describes): only recompute if the
    weights differ from it -- decided on the device, no host synchronisation."""
    nat.require_cuda(a, b, upstream)
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError(f"pair_fused expects two [N,C,H,W] tensors of equal shape, got {tuple(a.shape)} and {tuple(b.shape)}")
    if upstream.dtype != torch.float32 or upstream.numel() != nat.NLOSS:
        raise ValueError("upstream must be float32 [7]")
    a, a_sn, a_sc = nat.planes(a)
    b, b_sn, b_sc = nat.planes(b)
    n, c, h, w = a.shape
    L = nat.lib()
    nbytes = L.eco_pair_fused_ws_bytes(c)
    if nbytes < 0:
        raise ValueError(f"pair_fused serves at most 64 leaves per launch (got {c})")
    ws = nat.workspace("pairfused", nbytes, a.device)
    if upstream_prev is not None and (losses is None or sums is None or (want_a and out_a is None) or (want_b and out_b is None)):
        raise ValueError("upstream_prev needs the losses / sums / gradient buffers of the step it describes")
    if sums is None:
        sums = torch.empty((c, nat.NSTAT), dtype=torch.float64, device=a.device)
    if losses is None:
        losses = torch.empty((nat.NLOSS,), dtype=torch.float32, device=a.device)
    ga = (out_a if out_a is not None else torch.empty((n, c, h, w), dtype=a.dtype, device=a.device)) if want_a else None
    gb = (out_b if out_b is not None else torch.empty((n, c, h, w), dtype=b.dtype, device=b.device)) if want_b else None
    va, vb = nat.view_of(a, a_sn, a_sc), nat.view_of(b, b_sn, b_sc)
    oa, ob = nat.out_of(ga, c * h * w, h * w), nat.out_of(gb, c * h * w, h * w)
    rc = L.eco_pair_fused_ex(C.byref(va), C.byref(vb), n, c, h * w, int(flags), float(background_weight), float(scale),
                             _shape_ref(shape), upstream.data_ptr(),
                             upstream_prev.data_ptr() if upstream_prev is not None else None, ws.data_ptr(), ws.numel(),
                             sums.data_ptr(), losses.data_ptr(), C.byref(oa), C.byref(ob), _dev(a),
                             nat.current_stream_ptr(a.device))
    nat.check(rc, "eco_pair_fused")
    return losses, ga, gb, sums


def composite3_stats(x, g, from_logits):
    nat.require_cuda(x, g)
    if x.shape != g.shape or x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"composite3 expects two [N,3,H,W] tensors, got {tuple(x.shape)} and {tuple(g.shape)}")
    if g.dtype != torch.float32:
        g = g.float()
    x, x_sn, x_sc = nat.planes(x)
    g, g_sn, g_sc = nat.planes(g)
    n, _, h, w = x.shape
    L = nat.lib()
    ws = nat.workspace("comp3", L.eco_composite3_ws_bytes(), x.device)
    acc = torch.empty((nat.C3_NACC,), dtype=torch.float64, device=x.device)
    vx, vg = nat.view_of(x, x_sn, x_sc), nat.view_of(g, g_sn, g_sc)
    rc = L.eco_composite3_stats(C.byref(vx), C.byref(vg), n, h * w, int(from_logits), ws.data_ptr(), ws.numel(),
                                acc.data_ptr(), _dev(x), nat.current_stream_ptr(x.device))
    nat.check(rc, "eco_composite3_stats")
    return acc


def composite3_finalize(acc, leaf_scales, want_leaf_sums=False):
    dev = acc.device
    losses = torch.empty((nat.NLOSS,), dtype=torch.float32, device=dev)
    jac = torch.empty((nat.C3_NLEAF, nat.NLOSS, nat.NJAC), dtype=torch.float64, device=dev)
    leaf_sums = torch.empty((nat.C3_NLEAF, nat.NSTAT), dtype=torch.float64, device=dev) if want_leaf_sums else None
    if isinstance(leaf_scales, torch.Tensor):
        host, devp = None, leaf_scales.data_ptr()
    else:
        host, devp = (C.c_double * nat.C3_NLEAF)(*[float(s) for s in leaf_scales]), None
    rc = nat.lib().eco_composite3_finalize(acc.data_ptr(), host, devp, losses.data_ptr(), jac.data_ptr(),
                                           leaf_sums.data_ptr() if want_leaf_sums else None, _dev(acc),
                                           nat.current_stream_ptr(dev))
    nat.check(rc, "eco_composite3_finalize")
    return losses, jac, leaf_sums


def composite3_grad(x, g, from_logits, jac, upstream, out=None):
    if g.dtype != torch.float32:
        g = g.float()
    x, x_sn, x_sc = nat.planes(x)
    g, g_sn, g_sc = nat.planes(g)
    n, c, h, w = x.shape
    gx = out if out is not None else torch.empty((n, c, h, w), dtype=x.dtype, device=x.device)
    vx, vg = nat.view_of(x, x_sn, x_sc), nat.view_of(g, g_sn, g_sc)
    og = nat.out_of(gx, c * h * w, h * w)
    rc = nat.lib().eco_composite3_grad(C.byref(vx), C.byref(vg), n, h * w, int(from_logits), jac.data_ptr(),
                                       upstream.data_ptr(), C.byref(og), _dev(x), nat.current_stream_ptr(x.device))
    nat.check(rc, "eco_composite3_grad")
    return gx


def dice_counts(logits, labels, thresholds=None, inputs_are_probs=False, out_counts=None, out_soft=None):
    """One pass over logits+labels -> (counts int64 [T,C,3] = (I, |out|, |lab|) per threshold,
    soft float64 [C,3] = (sum p*lab, sum p, sum lab^2)).  thresholds: float32 CUDA tensor [T] or None.
    ``out_counts`` / ``out_soft``: optional contiguous destinations of those shapes (e.g. one slot of a stream buffer).
    Labels are expected to be 0/1 here (a class with other values gets |lab| = -1); ``dice_counts_ex`` serves any labels."""
    return dice_counts_ex(logits, labels, thresholds, inputs_are_probs, out_counts, out_soft, want_inter=False)[:2]


def dice_counts_ex(logits, labels, thresholds=None, inputs_are_probs=False, out_counts=None, out_soft=None, out_inter=None,
                   want_inter=True, ununion_preds=False):
    """``dice_counts`` plus inter float64 [T,C] = sum out*lab per threshold with the REAL label values (the integer
    intersection for 0/1 labels, an exact second pass over a class that holds anything else): what the reference's
    thresholded Dice (test_multiclass.py:80) needs when the dataset resized its masks.
    ``ununion_preds`` (soft Dice only): score ``return_union_sets_descending_order(sigmoid(logits), reverse=True)``
    (test_multiclass_sequential_densenetloss.py:66) -- the un-union happens in registers at load, ``logits`` stay as they are."""
    nat.require_cuda(logits, labels)
    if logits.shape != labels.shape or logits.dim() != 4:
        raise ValueError("dice_counts expects two [N,C,H,W] tensors of equal shape")
    if labels.dtype == torch.bool:
        labels = labels.view(torch.uint8)   # byte masks: 5 B/element instead of 8
    logits, z_sn, z_sc = nat.planes(logits)
    labels, l_sn, l_sc = nat.planes(labels)
    n, c, h, w = logits.shape
    nthr = 0 if thresholds is None else int(thresholds.numel())
    if nthr > 20:
        raise ValueError("at most 20 thresholds per call")
    if nthr:
        if thresholds.dtype != torch.float32 or not thresholds.is_cuda:
            raise ValueError("thresholds must be a float32 CUDA tensor")
        thresholds = thresholds.contiguous()
    L = nat.lib()
    ws = nat.workspace("dice%d" % nthr, L.eco_dice_ws_bytes(c, nthr), logits.device)
    if out_counts is None:
        counts = torch.zeros((max(nthr, 1), c, 3), dtype=torch.int64, device=logits.device)
    else:
        counts = out_counts
        if counts.dtype != torch.int64 or counts.numel() != max(nthr, 1) * c * 3 or not counts.is_contiguous():
            raise ValueError("out_counts must be a contiguous int64 tensor of max(T,1)*C*3 elements")
    if out_soft is None:
        soft = torch.empty((c, 3), dtype=torch.float64, device=logits.device)
    else:
        soft = out_soft
        if soft.dtype != torch.float64 or soft.numel() != c * 3 or not soft.is_contiguous():
            raise ValueError("out_soft must be a contiguous float64 tensor of C*3 elements")
    inter = None
    if want_inter and nthr:
        if out_inter is None:
            inter = torch.empty((nthr, c), dtype=torch.float64, device=logits.device)
        else:
            inter = out_inter
            if inter.dtype != torch.float64 or inter.numel() != nthr * c or not inter.is_contiguous():
                raise ValueError("out_inter must be a contiguous float64 tensor of T*C elements")
    vz, vl = nat.view_of(logits, z_sn, z_sc), nat.view_of(labels, l_sn, l_sc, allow_u8=True)
    rc = L.eco_dice_counts_ex(C.byref(vz), C.byref(vl), n, c, h * w, thresholds.data_ptr() if nthr else None, nthr,
                              (nat.EVAL_PROBS if inputs_are_probs else 0) | (nat.EVAL_UNUNION if ununion_preds else 0),
                              ws.data_ptr(), ws.numel(), counts.data_ptr(), soft.data_ptr(),
                              inter.data_ptr() if inter is not None else None, _dev(logits),
                              nat.current_stream_ptr(logits.device))
    nat.check(rc, "eco_dice_counts")
    return counts, soft, inter


def dice_finalize(counts, soft, nthr, inter=None):
    """counts / soft (/ inter, see dice_counts_ex) -> (thresholded Dice f32 [T,C] or None, soft Dice f32 [C])."""
    c = soft.shape[0]
    dev = soft.device
    dice = torch.empty((max(nthr, 1), c), dtype=torch.float32, device=dev) if nthr else None
    sdice = torch.empty((c,), dtype=torch.float32, device=dev)
    rc = nat.lib().eco_dice_finalize_ex(counts.data_ptr(), soft.data_ptr(), inter.data_ptr() if inter is not None else None,
                                        c, nthr, dice.data_ptr() if nthr else None, sdice.data_ptr(), _dev(soft),
                                        nat.current_stream_ptr(dev))
    nat.check(rc, "eco_dice_finalize")
    return dice, sdice


def masks_u8(x, threshold=None, inputs_are_probs=False):
    """[N,C,H,W] logits (or probabilities / labels) -> uint8 [N,C,H,W] = trunc(q * 255), q = sigmoid(x) (or x),
    thresholded first when ``threshold`` is given (test_multiclass.py:58,68-69,90-92): one 5 B/element pass."""
    nat.require_cuda(x)
    if x.dim() != 4:
        raise ValueError(f"masks_u8 expects a [N,C,H,W] tensor, got {tuple(x.shape)}")
    x, sn, sc = nat.planes(x)
    n, c, h, w = x.shape
    out = torch.empty((n, c, h, w), dtype=torch.uint8, device=x.device)
    v = nat.view_of(x, sn, sc)
    rc = nat.lib().eco_masks_u8(C.byref(v), n, c, h * w, float(threshold) if threshold is not None else 0.0,
                                int(threshold is not None), int(inputs_are_probs), out.data_ptr(), _dev(x),
                                nat.current_stream_ptr(x.device))
    nat.check(rc, "eco_masks_u8")
    return out


def softce_stats(a, b, need_bg):
    nat.require_cuda(a, b)
    a, a_sn, a_sc = nat.planes(a)
    b, b_sn, b_sc = nat.planes(b)
    n, c, h, w = a.shape
    L = nat.lib()
    ws = nat.workspace("softce", L.eco_softce_ws_bytes(), a.device)
    sums = torch.zeros((2,), dtype=torch.float64, device=a.device)
    va, vb = nat.view_of(a, a_sn, a_sc), nat.view_of(b, b_sn, b_sc)
    rc = L.eco_softce_stats(C.byref(va), C.byref(vb), n, c, h * w, int(need_bg), ws.data_ptr(), ws.numel(),
                            sums.data_ptr(), _dev(a), nat.current_stream_ptr(a.device))
    nat.check(rc, "eco_softce_stats")
    return sums


def softce_grad(a, b, bw, n_pix, upstream, want_a, want_b):
    a, a_sn, a_sc = nat.planes(a)
    b, b_sn, b_sc = nat.planes(b)
    n, c, h, w = a.shape
    ga = torch.empty((n, c, h, w), dtype=a.dtype, device=a.device) if want_a else None
    gb = torch.empty((n, c, h, w), dtype=b.dtype, device=b.device) if want_b else None
    va, vb = nat.view_of(a, a_sn, a_sc), nat.view_of(b, b_sn, b_sc)
    oa, ob = nat.out_of(ga, c * h * w, h * w), nat.out_of(gb, c * h * w, h * w)
    rc = nat.lib().eco_softce_grad(C.byref(va), C.byref(vb), n, c, h * w, float(bw), float(n_pix),
                                   upstream.data_ptr(), C.byref(oa), C.byref(ob), _dev(a),
                                   nat.current_stream_ptr(a.device))
    nat.check(rc, "eco_softce_grad")
    return ga, gb


# ------------------------------------------------------------------------------------------------
# autograd
# ------------------------------------------------------------------------------------------------
def _stack_upstream(grads, like):
    """7 optional 0-d grads -> one float32 [7] device tensor, without a host sync."""
    if all(gr is not None and gr.dtype == torch.float32 and gr.dim() == 0 for gr in grads):
        return torch.stack(grads)
    zero = None
    parts = []
    for gr in grads:
        if gr is None:
            if zero is None:
                zero = torch.zeros((), dtype=torch.float32, device=like.device)
            parts.append(zero)
        else:
            parts.append(gr.reshape(()).to(torch.float32))
    return torch.stack(parts)


class PairLeaves(torch.autograd.Function):
    """C independent 7-loss leaves over [N,C,H,W] slot tensors; returns the 7 totals over channels.

    ``group``: optional torch.distributed process group -- the batch is sharded across its ranks and the
    per-leaf sums are all-reduced before the closed forms."""

    @staticmethod
    def forward(ctx, a, b, background_weight, scale, flags, group, shape, key):
        # key: None, or a name for the call site -- a top-level losses_fn whose upstream weights repeat from step to step
        # (train_multiclass.py:145): the step then runs as ONE launch with the weights of that site's previous backward
        c = a.shape[1]
        ctx.set_materialize_grads(False)
        ctx.fast = (PLAIN_FAST_PATH and key is not None and ctx.needs_input_grad[1] and not ctx.needs_input_grad[0]
                    and _plain3_dropin_ok(a, b, background_weight, flags, group, shape))
        ctx.key = key
        if ctx.fast:
            # The live call of the reference's training loop (train_multiclass.py:139-141 with three organs: labels in the
            # gt slot, predictions in the pred slot): ONE launch of the fused plain step with the upstream weights anticipated
            # from the previous backward; the backward only confirms them (see Composite3).
            dev = b.device
            used = _anticipated_upstream.get((key, dev.index))
            if used is None:
                used = _zero_upstream(dev)
            probs = not (flags & nat.FLAG_B_LOGIT)
            losses, gb = multiclass3_fused(b.detach(), a.detach(), scale, used, probs=probs)
            ctx.save_for_backward(a, b, losses, gb)
            ctx.holds, ctx.returned, ctx.scale, ctx.probs = used, False, float(scale), probs
            return tuple(losses.clone().unbind(0))   # (the saved `losses` buffer is rewritten by the backward's launch)
        want_a, want_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if PLAIN_FAST_PATH and key is not None and group is None and c <= 64 and (want_a or want_b) and a.dim() == 4:
            # Any other top-level leaf call (one organ -- cfg1, ORGANS=whole_body: the prediction sits in the gt slot and
            # background_weight counts --, organ counts other than three, bf16, strided slices, both slots with gradient): the
            # one-launch step of eco_pair_fused, the same way.  (The stand-alone primitives and the inner leaves of the
            # pair-by-pair composite pass no key: their upstream weights differ from call to call.)
            dev = a.device
            used = _anticipated_upstream.get((key, dev.index))
            if used is None:
                used = _zero_upstream(dev)
            try:
                losses, ga, gb, sums = pair_fused(a.detach(), b.detach(), flags, background_weight, scale, used, want_a, want_b,
                                                  shape=shape)
            except nat.EcoLossError:
                losses = None   # (more leaves than one resident wave holds: the three launches below)
            if losses is not None:
                ctx.fast = 2
                ctx.save_for_backward(a, b, losses, sums, ga, gb)
                ctx.holds, ctx.returned = used, False
                ctx.call = (flags, float(background_weight), float(scale), shape)
                return tuple(losses.clone().unbind(0))
        sums = pair_stats(a.detach(), b.detach(), flags | (nat.FLAG_NEED_BG if background_weight != 0 else 0), shape)
        sums = dist_.allreduce_sums_(sums, group)
        _, total, jac = pair_finalize(sums, background_weight, [scale] * c, shape)
        ctx.save_for_backward(a, b, jac)
        ctx.flags, ctx.shape = flags, shape
        return tuple(total.unbind(0))

    @staticmethod
    def backward(ctx, *grads):
        want_a, want_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (want_a or want_b) or all(g is None for g in grads):
            return None, None, None, None, None, None, None, None
        if ctx.fast == 2:
            a, b, losses, sums, ga, gb = ctx.saved_tensors
            up = _stack_upstream(grads, a)
            _anticipated_upstream[(ctx.key, a.device.index)] = up
            flags, bw, scale, shape = ctx.call
            if ctx.returned:   # a second backward through the same graph: autograd may own the first buffers by now
                _, ga, gb, _ = pair_fused(a.detach(), b.detach(), flags, bw, scale, up, want_a, want_b, shape=shape)
                return ga, gb, None, None, None, None, None, None
            pair_fused(a.detach(), b.detach(), flags, bw, scale, up, want_a, want_b, out_a=ga, out_b=gb, shape=shape,
                       upstream_prev=ctx.holds, losses=losses, sums=sums)
            ctx.holds, ctx.returned = up, True
            return ga, gb, None, None, None, None, None, None
        if ctx.fast:
            a, b, losses, gb = ctx.saved_tensors
            up = _stack_upstream(grads, b)
            _anticipated_upstream[(ctx.key, b.device.index)] = up
            if ctx.returned:   # a second backward through the same graph: autograd may own the first buffer by now
                _, gb = multiclass3_fused(b.detach(), a.detach(), ctx.scale, up, probs=ctx.probs)
                return None, gb, None, None, None, None, None, None
            multiclass3_fused(b.detach(), a.detach(), ctx.scale, up, out=gb, probs=ctx.probs, upstream_prev=ctx.holds,
                              losses=losses)
            ctx.holds, ctx.returned = up, True
            return None, gb, None, None, None, None, None, None
        a, b, jac = ctx.saved_tensors
        up = _stack_upstream(grads, a)
        ga, gb = pair_grad(a.detach(), b.detach(), ctx.flags, jac, up, want_a, want_b, ctx.shape)
        return ga, gb, None, None, None, None, None, None


PLAIN_FAST_PATH = True   # (switch for measurements: False sends the plain leaf calls through the three pair-leaf launches)


def _plain3_dropin_ok(a, b, background_weight, flags, group, shape):
    """Calls the one-launch plain 3-organ step serves: three channels, labels (a) and fp32 predictions (b: probabilities, or
    logits with FLAG_B_LOGIT) contiguous with 16-byte aligned planes, the default leaf shape, no background term, one GPU."""
    if group is not None or shape is not None or background_weight != 0 or (flags & ~nat.FLAG_B_LOGIT):
        return False
    if b.dim() != 4 or b.shape[1] != 3 or a.shape != b.shape or b.dtype != torch.float32 or a.dtype != torch.float32:
        return False
    hw = b.shape[2] * b.shape[3]
    return (hw % 4 == 0 and a.is_contiguous() and b.is_contiguous() and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
            and b.shape[0] * hw <= 2 ** 31)


# Anticipated upstream weights of the drop-in composite path, per device: the float32 [7] vector dT/dloss_k that the LAST
# backward through Composite3 received.  train() forms T = focal_dice_w*focal_dice + bce_l_w*bce_l + generalized_dice_w*
# (generalized_dice + twersky_dice) with weights that depend on the epoch only (train_multiclass.py:92-100,145), so the
# next step's backward will almost always receive the same vector.  Tensors in here are never modified in place.
_anticipated_upstream = {}
_scales_cache = {}


def _device_scales(leaf_scales, device):
    """21 python floats -> float64 CUDA tensor, re-uploaded only when the values change (they do not unless
    early_stopped draws random weights, loss_composite.py:49-52)."""
    key = (device.index, tuple(float(v) for v in leaf_scales))
    hit = _scales_cache.get(device.index)
    if hit is not None and hit[0] == key:
        return hit[1]
    t = torch.tensor(key[1], dtype=torch.float64, device=device)
    _scales_cache[device.index] = (key, t)
    return t


def _fused_dropin_ok(x, g, from_logits, group):
    """Inputs the third-generation fused kernel serves (so that the backward's "only if changed" launch is valid too):
    fp32 logits or -- the reference's own call order, F.sigmoid before losses_fn (train_multiclass.py:134) -- fp32
    probabilities."""
    if group is not None or x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
        return False
    hw = x.shape[2] * x.shape[3]
    if hw % 4 or not x.is_contiguous() or x.data_ptr() % 16 or not g.is_contiguous() or g.data_ptr() % 16:
        return False
    if g.dtype == torch.float32:
        return True
    return g.dtype in (torch.uint8, torch.bool) and hw % 16 == 0


class Composite3(torch.autograd.Function):
    """Fused 21-leaf composite loss for C == 3 (loss_composite.py:21-94, composite_set_theory=True).

    Fast path (fp32 logits, one GPU): the forward is ONE launch of the fused step with the upstream weights anticipated
    from the previous backward; it already leaves d(sum_k w_k loss_k)/d logits in a buffer.  The backward launches the
    "only if changed" form of the same kernel with the weights autograd really delivered: the kernel compares the two
    vectors on the device and returns at once when they agree (no host synchronisation), otherwise it redoes the step.
    Everything else: statistics kernel -> (all-reduce) -> closed forms in the forward, gradient kernel in the backward."""

    @staticmethod
    def forward(ctx, x, g, leaf_scales, from_logits, group):
        ctx.from_logits = from_logits
        ctx.set_materialize_grads(False)
        ctx.fast = ctx.needs_input_grad[0] and not ctx.needs_input_grad[1] and _fused_dropin_ok(x, g, from_logits, group)
        if ctx.fast:
            dev = x.device
            used = _anticipated_upstream.get(dev.index)
            if used is None:
                used = torch.zeros(nat.NLOSS, dtype=torch.float32, device=dev)
            scales = _device_scales(leaf_scales, dev)
            losses, gx = _dropin_prepared(x, g, scales, used, from_logits).run(upstream=used, scales=scales)
            ctx.save_for_backward(x, g, scales, losses, gx)
            ctx.holds = used          # the weights `gx` currently holds the gradient for
            ctx.returned = False
            return tuple(losses.clone().unbind(0))   # (the saved `losses` buffer is rewritten by the backward's launch)
        if not ctx.needs_input_grad[0] and not ctx.needs_input_grad[1] and _fused_dropin_ok(x, g, from_logits, group):
            # loss values only (losses_fn under torch.no_grad(), train_multiclass.py:175-198): the statistics pass and the
            # closed forms of the same kernel, one launch, no gradient pass
            dev = x.device
            scales = _device_scales(leaf_scales, dev)
            zero = _zero_upstream(dev)
            losses, _ = _dropin_prepared(x, g, scales, zero, from_logits).run(upstream=zero, scales=scales, no_grad=True)
            return tuple(losses.unbind(0))
        acc = composite3_stats(x.detach(), g.detach(), from_logits)
        acc = dist_.allreduce_sums_(acc, group)
        losses, jac, _ = composite3_finalize(acc, leaf_scales)
        ctx.save_for_backward(x, g, jac)
        return tuple(losses.unbind(0))

    @staticmethod
    def backward(ctx, *grads):
        if ctx.needs_input_grad[1]:
            raise RuntimeError("the fused composite kernel produces gradients for the predictions only; "
                               "loss_composite.losses_fn routes labels that require grad to the pair-leaf kernels")
        if not ctx.needs_input_grad[0] or all(gr is None for gr in grads):
            return None, None, None, None, None
        if ctx.fast:
            x, g, scales, losses, gx = ctx.saved_tensors
            up = _stack_upstream(grads, x)
            _anticipated_upstream[x.device.index] = up
            ent = _dropin_prepared(x, g, scales, up, ctx.from_logits)
            if ctx.returned:   # a second backward through the same graph: autograd may own the first buffer by now
                _, gx = ent.run(upstream=up, scales=scales)
                return gx, None, None, None, None
            ent.run(out=gx, upstream=up, scales=scales, upstream_prev=ctx.holds, losses=losses)
            ctx.holds, ctx.returned = up, True
            return gx, None, None, None, None
        x, g, jac = ctx.saved_tensors
        up = _stack_upstream(grads, x)
        gx = composite3_grad(x.detach(), g.detach(), ctx.from_logits, jac, up)
        return gx, None, None, None, None


class SoftCE(torch.autograd.Function):
    """F.cross_entropy(b, a) + bw * F.cross_entropy(1-b, 1-a) with float targets (loss_functions.py:29-30)."""

    @staticmethod
    def forward(ctx, a, b, background_weight):
        sums = softce_stats(a.detach(), b.detach(), background_weight != 0)
        n_pix = a.shape[0] * a.shape[2] * a.shape[3]
        ctx.save_for_backward(a, b)
        ctx.bw, ctx.n_pix = float(background_weight), n_pix
        return (-(sums[0] + background_weight * sums[1]) / n_pix).to(torch.float32)

    @staticmethod
    def backward(ctx, grad):
        a, b = ctx.saved_tensors
        want_a, want_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        up = grad.reshape(1).to(torch.float32)
        ga, gb = softce_grad(a.detach(), b.detach(), ctx.bw, ctx.n_pix, up, want_a, want_b)
        return ga, gb, None


# ------------------------------------------------------------------------------------------------
# shape plumbing for the stand-alone primitives (reduce over the WHOLE tensor pair)
# ------------------------------------------------------------------------------------------------
def as_single_leaf(a, b):
    """Two equal-shape tensors -> [N,1,H,W]-shaped views describing ONE leaf over all their elements."""
    nat.require_cuda(a, b)
    if a.shape != b.shape:
        a, b = torch.broadcast_tensors(a, b)
    if a.dim() == 4 and a.shape[1] == 1:
        return a, b  # the reference's usual case: channel slices; planes() handles the strides
    return a.contiguous().view(1, 1, 1, -1), b.contiguous().view(1, 1, 1, -1)


def leaf7(a, b, background_weight=0.0, scale=1.0, flags=0, group=None, shape=None, key=None):
    """The 7 losses of ONE leaf over all elements of (a, b), each times ``scale``.  ``shape``: optional
    (focal_gamma, tversky_alpha, tversky_beta, focal_dice_gamma) when a primitive is called with its own keywords."""
    a4, b4 = as_single_leaf(a, b)
    return PairLeaves.apply(a4, b4, float(background_weight), float(scale), int(flags), group, shape, key)


def multiclass3_fused(x, g, leaf_scale, upstream, out=None, probs=False, upstream_prev=None, losses=None):
    """ONE cooperative launch for the plain 3-organ multi-class loss (train_multiclass.py:253-274, C == 3): losses of
    the three (g_c, sigmoid(x_c)) leaves summed over channels, each times ``leaf_scale``, and the gradient of
    sum_k upstream[k] * loss_k w.r.t. the LOGITS x.  Returns (losses f32 [7], grad).
    ``probs``: x already are probabilities (F.sigmoid at train_multiclass.py:134 ran before), gradient w.r.t. them.
    ``upstream_prev`` (+ ``out`` / ``losses`` holding that step): only recompute if the weights differ from it."""
    nat.require_cuda(x, g, upstream)
    if x.shape != g.shape or x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"multiclass3 expects two [N,3,H,W] tensors, got {tuple(x.shape)} and {tuple(g.shape)}")
    if upstream.dtype != torch.float32 or upstream.numel() != nat.NLOSS:
        raise ValueError("upstream must be float32 [7]")
    if g.dtype != torch.float32:
        g = g.float()
    x, x_sn, x_sc = nat.planes(x)
    g, g_sn, g_sc = nat.planes(g)
    n, c, h, w = x.shape
    L = nat.lib()
    ws = nat.workspace("mc3", L.eco_composite3_ws_bytes(), x.device)
    if upstream_prev is not None and (out is None or losses is None):
        raise ValueError("upstream_prev needs the `out` and `losses` buffers of the step it describes")
    if losses is None:
        losses = torch.empty((nat.NLOSS,), dtype=torch.float32, device=x.device)
    gx = out if out is not None else torch.empty((n, c, h, w), dtype=x.dtype, device=x.device)
    vx, vg = nat.view_of(x, x_sn, x_sc), nat.view_of(g, g_sn, g_sc)
    og = nat.out_of(gx, c * h * w, h * w)
    rc = L.eco_multiclass3_step(C.byref(vx), C.byref(vg), n, h * w, float(leaf_scale), upstream.data_ptr(),
                                upstream_prev.data_ptr() if upstream_prev is not None else None,
                                nat.C3_PROBS if probs else 0, ws.data_ptr(), ws.numel(), losses.data_ptr(), C.byref(og),
                                _dev(x), nat.current_stream_ptr(x.device))
    nat.check(rc, "eco_multiclass3_step")
    return losses, gx


def _byte_labels_ok(x, g, from_logits):
    """uint8 / bool masks go to the kernel as bytes when the byte-label kernel serves the case (fp32 inputs or bf16 logits,
    planes and tiles 16-byte aligned); otherwise they are widened to float32 here, as the reference does (train_multiclass.py:119-123)."""
    n, c, h, w = x.shape
    return ((x.dtype == torch.float32 or (from_logits and x.dtype == torch.bfloat16)) and (h * w) % 16 == 0
            and g.is_contiguous() and g.data_ptr() % 16 == 0)


def composite3_fused(x, g, leaf_scales, upstream, from_logits=True, out=None, union_labels=False, peers=None,
                     upstream_prev=None, losses_out=None):
    """ONE cooperative launch (eco_composite3_step): statistics -> grid hand-over -> closed forms -> gradient of
    sum_k upstream[k] * loss_k.  leaf_scales: float64 CUDA [21]; upstream: float32 CUDA [7].
    g: float32 labels, or uint8 / bool masks (kept as bytes on the device where the kernel takes them).
    union_labels: g holds the raw per-organ masks; the label union of utils/subsets_union.py:8-32 (exclude_indices=[0])
    is applied at load.  peers: an EcoPeerExchange for a batch sharded over processes.
    upstream_prev: `out` / `losses_out` already hold the step for these weights -- the launch returns at once unless
    `upstream` differs from them (compared on the device; eco_composite3_step_if_changed).
    Returns (losses f32 [7], grad w.r.t. x -- w.r.t. the logits when from_logits)."""
    nat.require_cuda(x, g, leaf_scales, upstream)
    if x.shape != g.shape or x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"composite3 expects two [N,3,H,W] tensors, got {tuple(x.shape)} and {tuple(g.shape)}")
    if leaf_scales.dtype != torch.float64 or leaf_scales.numel() != nat.C3_NLEAF:
        raise ValueError("leaf_scales must be float64 [21]")
    if upstream.dtype != torch.float32 or upstream.numel() != nat.NLOSS:
        raise ValueError("upstream must be float32 [7]")
    if g.dtype == torch.bool:
        g = g.view(torch.uint8)
    if g.dtype == torch.uint8:
        if not _byte_labels_ok(x, g, from_logits):
            g = g.float()
    elif g.dtype != torch.float32:
        g = g.float()
    x, x_sn, x_sc = nat.planes(x)
    g, g_sn, g_sc = nat.planes(g)
    n, c, h, w = x.shape
    L = nat.lib()
    ws = nat.workspace("comp3", L.eco_composite3_ws_bytes(), x.device)
    losses = losses_out if losses_out is not None else torch.empty((nat.NLOSS,), dtype=torch.float32, device=x.device)
    gx = out if out is not None else torch.empty((n, c, h, w), dtype=x.dtype, device=x.device)
    vx, vg = nat.view_of(x, x_sn, x_sc), nat.view_of(g, g_sn, g_sc, allow_u8=True)
    og = nat.out_of(gx, c * h * w, h * w)
    flags = (0 if from_logits else nat.C3_PROBS) | (nat.C3_UNION_LABELS if union_labels else 0)
    if upstream_prev is not None:
        rc = L.eco_composite3_step_if_changed(C.byref(vx), C.byref(vg), n, h * w, flags, leaf_scales.data_ptr(),
                                              upstream.data_ptr(), upstream_prev.data_ptr(), ws.data_ptr(), ws.numel(),
                                              losses.data_ptr(), C.byref(og), _dev(x), nat.current_stream_ptr(x.device))
        nat.check(rc, "eco_composite3_step_if_changed")
        return losses, gx
    rc = L.eco_composite3_step(C.byref(vx), C.byref(vg), n, h * w, flags, leaf_scales.data_ptr(), upstream.data_ptr(),
                               ws.data_ptr(), ws.numel(), losses.data_ptr(), C.byref(og),
                               C.byref(peers) if peers is not None else None, _dev(x), nat.current_stream_ptr(x.device))
    if rc == -8 and union_labels:
        # no fused union for this input (bf16 / probabilities / ragged planes): the in-place union kernel on a float copy
        from .subsets_union import return_union_sets_descending_order
        return composite3_fused(x, return_union_sets_descending_order(g.float().clone()), leaf_scales, upstream, from_logits,
                                out=out, union_labels=False, peers=peers)
    nat.check(rc, "eco_composite3_step")
    return losses, gx


class PreparedComposite3:
    """The argument block of one eco_composite3_step launch, built once for a (logits, labels) pair of buffers and reused:
    a training loop presents the same device buffers step after step, and re-deriving views, strides, workspace and
    ctypes structures costs more host time (~70 us) than the kernel takes.  ``run`` is a dictionary-free ctypes call."""

    __slots__ = ("sig", "vx", "vg", "og", "n", "hw", "flags", "scales", "upstream", "ws", "peers", "keep", "dev", "shape",
                 "dtype", "fn", "L")

    @staticmethod
    def signature(x, g):
        return (x.data_ptr(), g.data_ptr(), x.shape, g.shape, x.dtype, g.dtype, x.stride(), g.stride())

    def __init__(self, x, g, leaf_scales, upstream, from_logits=True, union_labels=False, peers=None):
        nat.require_cuda(x, g, leaf_scales, upstream)
        if x.shape != g.shape or x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"composite3 expects two [N,3,H,W] tensors, got {tuple(x.shape)} and {tuple(g.shape)}")
        self.sig = self.signature(x, g)
        gg = g.view(torch.uint8) if g.dtype == torch.bool else g
        if gg.dtype == torch.uint8 and not _byte_labels_ok(x, gg, from_logits):
            raise ValueError("byte labels not servable here")    # caller falls back to composite3_fused (which widens them)
        if gg.dtype not in (torch.uint8, torch.float32):
            raise ValueError("labels must be float32 / uint8 / bool")
        xx, x_sn, x_sc = nat.planes(x)
        g2, g_sn, g_sc = nat.planes(gg)
        if xx is not x or g2 is not gg:
            raise ValueError("strided planes")                    # a copy would not be reusable
        n, c, h, w = x.shape
        self.L = nat.lib()
        self.fn = self.L.eco_composite3_step
        self.n, self.hw, self.shape, self.dtype, self.dev = n, h * w, (n, c, h, w), x.dtype, x.device
        self.vx, self.vg = nat.view_of(x, x_sn, x_sc), nat.view_of(gg, g_sn, g_sc, allow_u8=True)
        self.og = nat.EcoOut(None, c * h * w, h * w, nat.dtype_code(x), 0)
        self.flags = (0 if from_logits else nat.C3_PROBS) | (nat.C3_UNION_LABELS if union_labels else 0)
        self.scales, self.upstream, self.peers = leaf_scales, upstream, peers
        self.ws = nat.workspace("comp3", self.L.eco_composite3_ws_bytes(), x.device)
        self.keep = (x, g, gg)

    def run(self, out=None, upstream=None, scales=None, upstream_prev=None, losses=None, no_grad=False):
        """Launch.  ``upstream`` / ``scales`` override the tensors given at construction; with ``upstream_prev`` the launch is
        the "only if changed" form (``out`` and ``losses`` then hold the step for ``upstream_prev``); ``no_grad``: loss values
        only (ECO_C3_NO_GRAD), the returned gradient is None."""
        dev = self.dev
        if losses is None:
            losses = torch.empty((nat.NLOSS,), dtype=torch.float32, device=dev)
        if no_grad:
            self.og.ptr = None
            rc = self.fn(self.vx, self.vg, self.n, self.hw, self.flags | nat.C3_NO_GRAD,
                         (self.scales if scales is None else scales).data_ptr(),
                         (self.upstream if upstream is None else upstream).data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                         losses.data_ptr(), self.og, self.peers, dev.index, torch._C._cuda_getCurrentRawStream(dev.index))
            if rc:
                nat.check(rc, "eco_composite3_step")
            return losses, None
        gx = out if out is not None else torch.empty(self.shape, dtype=self.dtype, device=dev)
        self.og.ptr = gx.data_ptr()
        up = (self.upstream if upstream is None else upstream).data_ptr()
        sc = (self.scales if scales is None else scales).data_ptr()
        stream = torch._C._cuda_getCurrentRawStream(dev.index)
        if upstream_prev is None:
            rc = self.fn(self.vx, self.vg, self.n, self.hw, self.flags, sc, up, self.ws.data_ptr(), self.ws.numel(),
                         losses.data_ptr(), self.og, self.peers, dev.index, stream)
        else:
            rc = self.L.eco_composite3_step_if_changed(self.vx, self.vg, self.n, self.hw, self.flags, sc, up,
                                                       upstream_prev.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                                                       losses.data_ptr(), self.og, dev.index, stream)
        if rc:
            nat.check(rc, "eco_composite3_step")
        return losses, gx


_dropin_launches = {}


def _dropin_prepared(x, g, scales, upstream, from_logits=True):
    """Launch block of the drop-in fast path for this pair of buffers (see PreparedComposite3)."""
    key = (x.data_ptr(), g.data_ptr(), torch._C._cuda_getCurrentRawStream(x.device.index), bool(from_logits))
    ent = _dropin_launches.get(key)
    if ent is None or ent.sig != PreparedComposite3.signature(x, g):
        if len(_dropin_launches) >= 64:
            _dropin_launches.clear()
        ent = _dropin_launches[key] = PreparedComposite3(x, g, scales, upstream, bool(from_logits))
    return ent


_zero_up = {}


def _zero_upstream(device):
    t = _zero_up.get(device.index)
    if t is None:
        t = _zero_up[device.index] = torch.zeros(nat.NLOSS, dtype=torch.float32, device=device)
    return t
