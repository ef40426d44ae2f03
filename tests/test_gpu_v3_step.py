"""The third-generation fused composite step (eco_composite3_step): byte labels, the label union folded into the load
stage, the in-kernel peer exchange exercised by two ranks on ONE device, and its time-out reporting.
Reference lines: ess/loss_composite.py:21-94, ess/train_multiclass.py:110,119-123,133-147, ess/utils/subsets_union.py:8-32."""
import ctypes as C

import numpy as np
import pytest
import torch

from parity import assert_grad_close, assert_losses_close, TOL

pytestmark = pytest.mark.gpu

UP = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]
UP_ALL = [0.3, 1.0, 0.7, 0.2, 1.0, 0.5, 1.0]


def _oracle(z, g, up):
    from oracle import torch_port as tp
    zr = z.clone().requires_grad_(True)
    np.random.seed(0)
    ref = tp.losses_composite(torch.sigmoid(zr), g, True)
    sum(w * l for w, l in zip(up, ref) if w).backward()
    return [float(v) for v in ref], zr.grad


@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (3, 3, 48, 48), (5, 3, 64, 64), (54, 3, 256, 256)])
@pytest.mark.parametrize("up", [UP, UP_ALL])
def test_v3_vs_oracle_and_byte_labels_bit_identical(shape, up):
    """fp32 labels against the same-device oracle; uint8 and bool masks must give bit-identical results (the bytes are
    widened exactly in registers, the arithmetic is the same)."""
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    z, g = make_inputs(shape[0], 3, shape[2], 7, nested=True)
    z, g = z.cuda(), g.cuda()
    np.random.seed(0)
    step = CompositeLossStep(up)
    l32, d32 = step(z, g)
    rl, rg = _oracle(z, g, up)
    assert_losses_close(l32.cpu().numpy(), rl, tol=TOL, what=f"v3 {shape}")
    assert_grad_close(d32.cpu(), rg.cpu(), tol=TOL, what=f"v3 {shape}")
    l8, d8 = step(z, g.to(torch.uint8))
    assert torch.equal(l8, l32) and torch.equal(d8, d32)
    lb, db = step(z, g.bool())
    assert torch.equal(lb, l32) and torch.equal(db, d32)


def test_v3_is_deterministic_and_rearms_its_workspace():
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    z, g = make_inputs(9, 3, 128, 11)
    z, g = z.cuda(), g.cuda()
    step = CompositeLossStep(UP)
    first = step(z, g)
    for _ in range(5):   # odd and even step parities
        l, d = step(z, g)
        assert torch.equal(l, first[0]) and torch.equal(d, first[1])


def test_v3_label_union_folded_into_the_load_stage():
    """raw disjoint masks + ECO_C3_UNION_LABELS == explicit in-place union (utils/subsets_union.py:8-32) + plain step,
    bit for bit, for float and byte masks; the caller's masks are not modified."""
    from ecologysemanticsegmentation_b200 import subsets_union as su
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    z, g = make_inputs(6, 3, 64, 21)
    z = z.cuda()
    raw = torch.stack([g[:, 0], g[:, 1] - g[:, 2], g[:, 2]], 1).contiguous().cuda()   # whole_body, ventral only, dorsal
    raw[0, 1, :4, :4] = 1.0
    raw[0, 2, :4, :4] = 1.0            # overlapping pixels: the sum 2 must be clamped to 1
    raw[1, 0, 0, :8] = 3.0             # and everything above 1 anywhere (:28 of the reference)
    united = su.return_union_sets_descending_order(raw.clone())
    ref_l, ref_d = CompositeLossStep(UP)(z, united)
    keep = raw.clone()
    l, d = CompositeLossStep(UP, union_labels=True)(z, raw)
    assert torch.equal(raw, keep)
    assert torch.equal(l, ref_l) and torch.equal(d, ref_d)
    raw8 = torch.stack([g[:, 0], g[:, 1] - g[:, 2], g[:, 2]], 1).contiguous().to(torch.uint8).cuda()
    united8 = su.return_union_sets_descending_order(raw8.float())
    ref_l, ref_d = CompositeLossStep(UP)(z, united8)
    l, d = CompositeLossStep(UP, union_labels=True)(z, raw8)
    assert torch.equal(l, ref_l) and torch.equal(d, ref_d)
    # inputs the fused union does not serve (bf16 logits) fall back to the explicit union
    l16, d16 = CompositeLossStep(UP, union_labels=True)(z.bfloat16(), raw)
    r16, rd16 = CompositeLossStep(UP)(z.bfloat16(), united)
    assert torch.equal(l16, r16) and torch.equal(d16, rd16)


def test_v3_nonbinary_labels_take_the_exact_slow_path():
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    torch.manual_seed(11)
    z = torch.randn(3, 3, 32, 32).cuda()
    g = (torch.rand(3, 3, 32, 32) > 0.5).float()
    g[0, 1, 3, 4] = 0.25
    g[1, 2, 0, 0] = 0.6
    g[2, 0, 5, 5] = 0.5
    g = g.cuda()
    np.random.seed(0)
    l, d = CompositeLossStep(UP_ALL)(z, g)
    rl, rg = _oracle(z, g, UP_ALL)
    assert_losses_close(l.cpu().numpy(), rl, tol=TOL, what="v3 nonbinary")
    assert_grad_close(d.cpu(), rg.cpu(), tol=TOL, what="v3 nonbinary")


def _raw_step(L, nat, z, g, scales, up, ws, losses, gx, peers, stream):
    n, c, h, w = z.shape
    vx, vg = nat.view_of(z, c * h * w, h * w), nat.view_of(g, c * h * w, h * w, allow_u8=True)
    og = nat.out_of(gx, c * h * w, h * w)
    rc = L.eco_composite3_step(C.byref(vx), C.byref(vg), n, h * w, 0, scales.data_ptr(), up.data_ptr(), ws.data_ptr(), ws.numel(),
                               losses.data_ptr(), C.byref(og), C.byref(peers) if peers is not None else None,
                               z.device.index, stream.cuda_stream)
    nat.check(rc, "eco_composite3_step")


def test_peer_exchange_two_ranks_on_one_device():
    """The in-kernel all-reduce (LL stores into every rank's exchange buffer, every CTA receives) with TWO ranks as two
    concurrent cooperative launches on two streams of ONE device -- each rank's grid is small enough for both to be
    co-resident -- so the protocol is covered on a single-GPU box too.  The sharded result must equal the single-launch
    step on the concatenated batch: losses to 1e-6, each shard's gradient to 1e-6 of the max."""
    from ecologysemanticsegmentation_b200 import _native as nat
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    L = nat.lib()
    dev = torch.device("cuda", 0)
    z, g = make_inputs(8, 3, 64, 99)      # 2 ranks x 4 images x 4 tiles = 16 CTAs per rank
    z, g = z.cuda(), g.cuda()
    np.random.seed(0)
    single = CompositeLossStep(UP, device=dev)
    ref_l, ref_d = single(z, g)
    world = 2
    bufs = [torch.zeros(int(L.eco_xch_bytes(world)), dtype=torch.uint8, device=dev) for _ in range(world)]
    peer_ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32).pin_memory()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    wss = [torch.zeros(int(L.eco_composite3_ws_bytes()), dtype=torch.uint8, device=dev) for _ in range(world)]
    losses = [torch.empty(7, dtype=torch.float32, device=dev) for _ in range(world)]
    shards = [(z[:4].contiguous(), g[:4].contiguous()), (z[4:].contiguous(), g[4:].to(torch.uint8).contiguous())]  # rank 1: byte labels
    grads = [torch.empty_like(s[0]) for s in shards]
    torch.cuda.synchronize()
    for epoch in range(1, 6):
        for r in range(world):
            peers = nat.EcoPeerExchange(peer_ptrs.data_ptr(), r, world, epoch, 0, status.data_ptr(), 3000.0)
            _raw_step(L, nat, shards[r][0], shards[r][1], single.scales, single.upstream, wss[r], losses[r], grads[r], peers, streams[r])
        torch.cuda.synchronize()
        assert int(status[0]) == 0, "a rank timed out waiting for its peer (the two launches did not run concurrently?)"
        for r in range(world):
            rel = ((losses[r][1:] - ref_l[1:]).abs() / ref_l[1:].abs()).max()
            assert float(rel) <= 1e-6, (epoch, r, losses[r], ref_l)
            err = (grads[r] - ref_d[4 * r:4 * r + 4]).abs().max() / ref_d.abs().max()
            assert float(err) <= 1e-6, (epoch, r, float(err))
        assert torch.equal(losses[0], losses[1]), "every rank must hold bit-identical totals"


def test_peer_exchange_timeout_is_reported():
    """A peer that never shows up: the wait gives up after timeout_ms, the outputs are poisoned with NaN (not silently
    wrong) and the status word is set, where the host sees it without a synchronisation of its own."""
    from ecologysemanticsegmentation_b200 import _native as nat
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    L = nat.lib()
    dev = torch.device("cuda", 0)
    z, g = make_inputs(2, 3, 32, 5)
    z, g = z.cuda(), g.cuda()
    single = CompositeLossStep(UP, device=dev)
    world = 2
    bufs = [torch.zeros(int(L.eco_xch_bytes(world)), dtype=torch.uint8, device=dev) for _ in range(world)]
    peer_ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32).pin_memory()
    ws = torch.zeros(int(L.eco_composite3_ws_bytes()), dtype=torch.uint8, device=dev)
    losses = torch.zeros(7, dtype=torch.float32, device=dev)
    gx = torch.zeros_like(z)
    peers = nat.EcoPeerExchange(peer_ptrs.data_ptr(), 0, world, 1, 0, status.data_ptr(), 50.0)   # rank 1 never runs
    _raw_step(L, nat, z, g, single.scales, single.upstream, ws, losses, gx, peers, torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert int(status[0]) == 1
    assert not bool(torch.isfinite(losses[1:]).any()) and not bool(torch.isfinite(gx).all())
    # same through the workspace word + eco_xch_poll_status (status == NULL)
    peers = nat.EcoPeerExchange(peer_ptrs.data_ptr(), 0, world, 2, 0, None, 50.0)
    ws2 = torch.zeros(int(L.eco_composite3_ws_bytes()), dtype=torch.uint8, device=dev)
    _raw_step(L, nat, z, g, single.scales, single.upstream, ws2, losses, gx, peers, torch.cuda.current_stream())
    out = C.c_uint32(7)
    nat.check(L.eco_xch_poll_status(ws2.data_ptr(), ws2.numel(), C.byref(out), 0, torch.cuda.current_stream().cuda_stream), "poll")
    assert out.value == 1
    nat.check(L.eco_xch_poll_status(ws2.data_ptr(), ws2.numel(), C.byref(out), 0, torch.cuda.current_stream().cuda_stream), "poll")
    assert out.value == 0, "the word is cleared once it has been read"


def test_dropin_autograd_fast_path_anticipates_the_weights():
    """eco.losses_fn(z, g, True, from_logits=True) -> weighted sum -> backward(), as ess/train_multiclass.py:139-147 calls
    it: the forward is the fused step with the weights of the PREVIOUS backward anticipated, the backward an
    "only if changed" launch.  Every combination must give the oracle's gradient: first call (nothing anticipated), repeated
    weights (hit), changed weights (miss), two forwards before two backwards, and a second backward through one graph."""
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import ops
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    ops._anticipated_upstream.clear()
    z0, g0 = make_inputs(6, 3, 64, 31)
    z0, g0 = z0.cuda(), g0.cuda()

    def ours(z, g, up, retain=False):
        zz = z.clone().requires_grad_(True)
        np.random.seed(0)
        losses = eco.losses_fn(zz, g, True, from_logits=True)
        return zz, losses, sum(w * l for w, l in zip(up, losses) if w)

    def check(z, g, up):
        zz, losses, total = ours(z, g, up)
        total.backward()
        rl, rg = _oracle(z, g.float(), up)
        assert_losses_close([float(v) for v in losses], rl, tol=TOL, what=f"drop-in {up}")
        assert_grad_close(zz.grad.cpu(), rg.cpu(), tol=TOL, what=f"drop-in {up}")

    check(z0, g0, UP)            # nothing anticipated yet: the backward recomputes
    check(z0, g0, UP)            # hit
    check(z0 * 0.5, g0, UP)      # hit on other data
    check(z0, g0, UP_ALL)        # the weights changed: miss, recomputed
    check(z0, g0.to(torch.uint8), UP_ALL)   # byte masks through the reference signature
    # two forwards, then two backwards with different weights
    za, la, ta = ours(z0, g0, UP)
    zb, lb, tb = ours(z0 * 2.0, g0, UP_ALL)
    tb.backward()
    ta.backward()
    assert_grad_close(za.grad.cpu(), _oracle(z0, g0, UP)[1].cpu(), tol=TOL, what="interleaved a")
    assert_grad_close(zb.grad.cpu(), _oracle(z0 * 2.0, g0, UP_ALL)[1].cpu(), tol=TOL, what="interleaved b")
    # a second backward through the same graph with other weights
    zc, lc, _ = ours(z0, g0, UP)
    sum(w * l for w, l in zip(UP, lc) if w).backward(retain_graph=True)
    first = zc.grad.clone()
    zc.grad = None
    sum(w * l for w, l in zip(UP_ALL, lc) if w).backward()
    assert_grad_close(first.cpu(), _oracle(z0, g0, UP)[1].cpu(), tol=TOL, what="retain 1")
    assert_grad_close(zc.grad.cpu(), _oracle(z0, g0, UP_ALL)[1].cpu(), tol=TOL, what="retain 2")


@pytest.mark.parametrize("shape", [(3, 3, 32, 40), (54, 3, 256, 256), (2, 3, 6, 6)])
def test_v3_bf16_logits_within_1e2(shape):
    """north_star: 1e-2 in bf16.  bf16 logits (bf16 gradient out) go through the same fused kernel (H*W % 8 == 0; the
    6x6 case falls back to the first-generation kernels); byte masks give bit-identical results there too."""
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from parity import TOL_BF16
    torch.manual_seed(shape[0] * 100 + shape[2])
    z16 = torch.randn(shape).to(torch.bfloat16).cuda()
    u = torch.rand(shape[0], 1, shape[2], shape[3])
    g = torch.cat([(u < 0.5).float(), (u < 0.22).float(), (u < 0.11).float()], 1).cuda()
    np.random.seed(0)
    step = CompositeLossStep(UP)
    l, d = step(z16, g)
    assert d.dtype == torch.bfloat16
    rl, rg = _oracle(z16.float(), g, UP)
    assert_losses_close(l.cpu().numpy(), rl, tol=TOL_BF16, what=f"bf16 {shape}")
    dd = (d.float() - rg).double()
    assert float(dd.abs().max() / rg.abs().max()) <= TOL_BF16 and float(dd.norm() / rg.double().norm()) <= TOL_BF16
    if (shape[2] * shape[3]) % 16 == 0:
        l8, d8 = step(z16, g.to(torch.uint8))
        assert torch.equal(l8, l) and torch.equal(d8, d)


def _oracle_probs(p, g, up):
    """gradient w.r.t. the PROBABILITIES, as autograd delivers it to F.sigmoid's backward (train_multiclass.py:134)"""
    from oracle import torch_port as tp
    pr = p.clone().requires_grad_(True)
    np.random.seed(0)
    ref = tp.losses_composite(pr, g, True)
    sum(w * l for w, l in zip(up, ref) if w).backward()
    return [float(v) for v in ref], pr.grad


@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (5, 3, 64, 64), (54, 3, 256, 256)])
@pytest.mark.parametrize("up", [UP, UP_ALL])
def test_v3_probability_inputs_one_launch(shape, up):
    """ECO_C3_PROBS on the one-launch kernel: the reference's own call order (F.sigmoid at train_multiclass.py:134, then
    losses_fn on the probabilities); losses and d/d probabilities against the same-device oracle, byte masks bit-identical.
    Exactly tied probabilities (|x_i - x_j| = 0: torch's abs backward is 0 there) are planted on purpose."""
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    z, g = make_inputs(shape[0], 3, shape[2], 13, nested=True)
    p = torch.sigmoid(z.cuda())
    p[0, 1, 0, :5] = p[0, 0, 0, :5]
    p[1, 2, 1, :3] = p[1, 1, 1, :3]
    g = g.cuda()
    np.random.seed(0)
    step = CompositeLossStep(up, from_logits=False)
    l32, d32 = step(p, g)
    rl, rg = _oracle_probs(p, g, up)
    assert_losses_close(l32.cpu().numpy(), rl, tol=TOL, what=f"v3 probs {shape}")
    assert_grad_close(d32.cpu(), rg.cpu(), tol=TOL, what=f"v3 probs {shape}")
    l8, d8 = step(p, g.to(torch.uint8))
    assert torch.equal(l8, l32) and torch.equal(d8, d32)


@pytest.mark.parametrize("from_logits", [True, False])
def test_v3_no_grad_step_equals_the_full_step(from_logits):
    """ECO_C3_NO_GRAD (losses_fn under torch.no_grad(), the validation loop of train_multiclass.py:175-198): the same
    statistics and closed forms without the gradient pass -- bit-identical loss values, several steps in a row (the
    workspace parity advances the same way), mixed with full steps."""
    from ecologysemanticsegmentation_b200 import ops
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    z, g = make_inputs(7, 3, 96, 17)
    x = z.cuda() if from_logits else torch.sigmoid(z.cuda())
    g = g.cuda()
    np.random.seed(0)
    step = CompositeLossStep(UP_ALL, from_logits=from_logits)
    full_l, full_d = step(x, g)
    ent = ops.PreparedComposite3(x, g, step.scales, step.upstream, from_logits)
    for _ in range(3):
        l, d = ent.run(no_grad=True)
        assert d is None and torch.equal(l, full_l)
        l2, d2 = ent.run()
        assert torch.equal(l2, full_l) and torch.equal(d2, full_d)


def test_dropin_unchanged_train_loop_probabilities_and_no_grad():
    """The reference's training step verbatim (train_multiclass.py:133-147): outputs = F.sigmoid(net(x)); losses =
    losses_fn(outputs, labels, True); weighted sum; backward -- on the one-launch kernel (gradient w.r.t. the probabilities,
    torch's sigmoid backward takes it to the logits), and the same call under torch.no_grad()."""
    from ecologysemanticsegmentation_b200 import loss_composite as lc
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    z, g = make_inputs(6, 3, 64, 5, nested=True)
    z, g = z.cuda(), g.cuda()
    rl, rg = _oracle(z, g, UP)
    for _ in range(3):   # first step: weights not anticipated yet; later steps: one launch + the no-op check
        zz = z.clone().requires_grad_(True)
        np.random.seed(0)
        losses = lc.losses_fn(torch.sigmoid(zz), g, True)
        sum(w * l for w, l in zip(UP, losses) if w).backward()
        assert_losses_close([float(v) for v in losses], rl, tol=TOL, what="drop-in probs")
        assert_grad_close(zz.grad.cpu(), rg.cpu(), tol=TOL, what="drop-in probs")
    with torch.no_grad():
        np.random.seed(0)
        val = lc.losses_fn(torch.sigmoid(z), g, True)
    assert_losses_close([float(v) for v in val], rl, tol=TOL, what="drop-in no_grad")
    with torch.no_grad():
        np.random.seed(0)
        val = lc.losses_fn(z, g, True, from_logits=True)
    assert_losses_close([float(v) for v in val], rl, tol=TOL, what="drop-in no_grad from logits")


def _oracle_plain(z, g, up, fn="losses_train_multiclass"):
    from oracle import torch_port as tp
    zr = z.clone().requires_grad_(True)
    ref = getattr(tp, fn)(torch.sigmoid(zr), g, False, 0, False)
    sum(w * l for w, l in zip(up, ref) if w).backward()
    return [float(v) for v in ref], zr.grad


def test_dropin_live_train_loop_plain_three_organs_one_launch():
    """The call the reference's train() really makes (ess/train_multiclass.py:134,139-141,145,147): outputs = F.sigmoid(net(x));
    losses_fn(outputs, labels, composite_set_theory=False, ...) with three organs; weighted sum; backward -- on ONE launch of
    the plain fused step on probabilities with anticipated weights.  First call, hit, miss, other data, interleaved forwards,
    a second backward through one graph, the doubled flavour of loss_composite.losses_fn, logits in (from_logits=True), and
    the three-launch path (labels that require grad, unaligned shapes) next to it."""
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import ops, train_multiclass as tm
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    ops._anticipated_upstream.clear()
    z0, g0 = make_inputs(6, 3, 64, 77)
    z0, g0 = z0.cuda(), g0.cuda()

    def ours(z, g, up, fn=tm.losses_fn, **kw):
        zz = z.clone().requires_grad_(True)
        x = zz if kw.get("from_logits") else torch.sigmoid(zz)
        losses = fn(x, g, False, 0, False, **kw)
        return zz, losses, sum(w * l for w, l in zip(up, losses) if w)

    def check(z, g, up, fn=tm.losses_fn, ref="losses_train_multiclass", **kw):
        zz, losses, total = ours(z, g, up, fn, **kw)
        total.backward()
        rl, rg = _oracle_plain(z, g, up, ref)
        assert_losses_close([float(v) for v in losses], rl, tol=TOL, what=f"live plain {up}")
        assert_grad_close(zz.grad.cpu(), rg.cpu(), tol=TOL, what=f"live plain {up}")

    check(z0, g0, UP)                       # nothing anticipated yet: the backward recomputes
    check(z0, g0, UP)                       # hit
    check(z0 * 0.5, g0, UP)                 # hit on other data
    check(z0, g0, UP_ALL)                   # the weights changed: miss, recomputed
    check(z0, g0, UP_ALL, from_logits=True)
    check(z0, g0, UP, fn=eco.losses_fn, ref="losses_composite")     # loss_composite.losses_fn: every leaf doubled
    za, la, ta = ours(z0, g0, UP)
    zb, lb, tb = ours(z0 * 2.0, g0, UP_ALL)
    tb.backward()
    ta.backward()
    assert_grad_close(za.grad.cpu(), _oracle_plain(z0, g0, UP)[1].cpu(), tol=TOL, what="interleaved a")
    assert_grad_close(zb.grad.cpu(), _oracle_plain(z0 * 2.0, g0, UP_ALL)[1].cpu(), tol=TOL, what="interleaved b")
    zc, lc, _ = ours(z0, g0, UP)
    sum(w * l for w, l in zip(UP, lc) if w).backward(retain_graph=True)
    first = zc.grad.clone()
    zc.grad = None
    sum(w * l for w, l in zip(UP_ALL, lc) if w).backward()
    assert_grad_close(first.cpu(), _oracle_plain(z0, g0, UP)[1].cpu(), tol=TOL, what="retain 1")
    assert_grad_close(zc.grad.cpu(), _oracle_plain(z0, g0, UP_ALL)[1].cpu(), tol=TOL, what="retain 2")
    # the same call with the fast path switched off gives the same numbers (three pair-leaf launches)
    ops.PLAIN_FAST_PATH = False
    try:
        zs, ls, ts = ours(z0, g0, UP)
        ts.backward()
    finally:
        ops.PLAIN_FAST_PATH = True
    zf, lf, tf = ours(z0, g0, UP)
    tf.backward()
    assert_losses_close([float(v) for v in lf], [float(v) for v in ls], tol=1e-6, what="fast vs three-launch")
    assert_grad_close(zf.grad.cpu(), zs.grad.cpu(), tol=1e-6, what="fast vs three-launch")
    # full cfg2 shape
    z2, g2 = make_inputs(54, 3, 256, 102)
    check(z2.cuda(), g2.cuda(), UP)


@pytest.mark.parametrize("shape", [(6, 1, 64, 64), (54, 1, 256, 256), (3, 1, 17, 19), (4, 2, 32, 32), (2, 5, 8, 24)])
def test_dropin_live_train_loop_other_organ_counts_one_launch(shape):
    """The same for every other organ count -- cfg1 (ORGANS=whole_body, the reference's default: C == 1, prediction in the gt
    slot, background_weight honoured) and plain C != 3 -- on the one-launch leaf step (eco_pair_fused_ex) with anticipated
    weights: first call, hit, miss, interleaved forwards, second backward through one graph, both losses_fn flavours."""
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import ops, train_multiclass as tm
    from oracle import torch_port as tp
    ops._anticipated_upstream.clear()
    n, c, h, w = shape
    torch.manual_seed(500 + c + h)
    z0 = torch.randn(n, c, h, w).cuda()
    g0 = (torch.rand(n, c, h, w) > 0.5).float().cuda()
    bw = 0.5 if c == 1 else 0

    def oracle(z, up, fn):
        zr = z.clone().requires_grad_(True)
        ref = getattr(tp, fn)(torch.sigmoid(zr), g0, False, bw, False)
        sum(wk * l for wk, l in zip(up, ref) if wk).backward()
        return [float(v) for v in ref], zr.grad

    def ours(z, up, fn):
        zz = z.clone().requires_grad_(True)
        losses = fn(torch.sigmoid(zz), g0, False, bw, False)
        return zz, losses, sum(wk * l for wk, l in zip(up, losses) if wk)

    def check(z, up, fn=tm.losses_fn, ref="losses_train_multiclass"):
        zz, losses, total = ours(z, up, fn)
        total.backward()
        rl, rg = oracle(z, up, ref)
        assert_losses_close([float(v) for v in losses], rl, tol=TOL, what=f"live {shape} {up}")
        assert_grad_close(zz.grad.cpu(), rg.cpu(), tol=TOL, what=f"live {shape} {up}")

    check(z0, UP)
    check(z0, UP)
    check(z0 * 0.5, UP)
    check(z0, UP_ALL)
    check(z0, UP, fn=eco.losses_fn, ref="losses_composite")
    check(z0, UP, fn=eco.losses_fn, ref="losses_composite")
    za, la, ta = ours(z0, UP, tm.losses_fn)
    zb, lb, tb = ours(z0 * 2.0, UP_ALL, tm.losses_fn)
    tb.backward()
    ta.backward()
    assert_grad_close(za.grad.cpu(), oracle(z0, UP, "losses_train_multiclass")[1].cpu(), tol=TOL, what="interleaved a")
    assert_grad_close(zb.grad.cpu(), oracle(z0 * 2.0, UP_ALL, "losses_train_multiclass")[1].cpu(), tol=TOL, what="interleaved b")
    zc, lc_, _ = ours(z0, UP, tm.losses_fn)
    sum(wk * l for wk, l in zip(UP, lc_) if wk).backward(retain_graph=True)
    first = zc.grad.clone()
    zc.grad = None
    sum(wk * l for wk, l in zip(UP_ALL, lc_) if wk).backward()
    assert_grad_close(first.cpu(), oracle(z0, UP, "losses_train_multiclass")[1].cpu(), tol=TOL, what="retain 1")
    assert_grad_close(zc.grad.cpu(), oracle(z0, UP_ALL, "losses_train_multiclass")[1].cpu(), tol=TOL, what="retain 2")
