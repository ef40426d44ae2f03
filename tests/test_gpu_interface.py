"""Behaviour of the drop-in interface on the GPU: layouts, dtypes, edge cases, the scoring loop."""
import numpy as np
import pytest
import torch

from parity import assert_grad_close, assert_losses_close, TOL, TOL_BF16

pytestmark = pytest.mark.gpu
UP = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]
UP_ALL = [0.3, 1.0, 0.7, 0.2, 1.0, 0.5, 1.0]


def _combine(losses, up):
    return sum(float(w) * l for w, l in zip(up, losses) if w != 0.0)


def _ref(fn_name, x, g, up, *args):
    from oracle import torch_port as tp
    x = x.detach().clone().requires_grad_(True)
    losses = getattr(tp, fn_name)(x, g, *args)
    _combine(losses, up).backward()
    return [float(v) for v in losses], x.grad


@pytest.mark.parametrize("shape", [(1, 3, 1, 1), (2, 3, 5, 3), (1, 3, 7, 9), (3, 3, 1, 130)])
def test_ragged_and_tiny_shapes_composite(shape):
    """H*W not a multiple of 4 -> scalar path; single pixel; one image."""
    import ecologysemanticsegmentation_b200 as eco
    torch.manual_seed(shape[2] * 100 + shape[3])
    p = (torch.rand(shape) * 0.9 + 0.05).cuda()
    g = (torch.rand(shape) > 0.5).float().cuda()
    np.random.seed(0)
    rl, rg = _ref("losses_composite", p, g, UP_ALL, True)
    x = p.clone().requires_grad_(True)
    np.random.seed(0)
    ours = eco.losses_fn(x, g, True)
    _combine(ours, UP_ALL).backward()
    assert_losses_close(ours, rl, what=f"ragged {shape}")
    assert_grad_close(x.grad.cpu(), rg.cpu(), what=f"ragged {shape}")


def test_noncontiguous_channel_slices_and_offsets():
    """The reference slices x[:, c:c+1]; odd storage offsets force the unaligned scalar path."""
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import loss_functions as lf
    from oracle import torch_port as tp
    torch.manual_seed(9)
    big = (torch.rand(3, 5, 12, 10) * 0.9 + 0.05).cuda()
    lab = (torch.rand(3, 5, 12, 10) > 0.5).float().cuda()
    for c in (0, 3):
        xs, gs = big[:, c:c + 1], lab[:, c:c + 1]
        assert not xs.is_contiguous()
        ours = eco.losses_fn(xs, gs, False, 0.5)
        ref = tp.losses_composite(xs, gs, False, 0.5)
        assert_losses_close(ours, ref, what="channel slice")
        assert_losses_close([lf.dice_loss(gs, xs)], [tp.pair_dice(gs, xs)])
    flat = torch.rand(1001, device="cuda")
    a, b = flat[1:1000].view(1, 1, 27, 37), flat[2:1001].view(1, 1, 27, 37)   # 4-byte aligned only
    assert_losses_close(eco.losses_fn(a, b), tp.losses_composite(a, b), what="unaligned views")
    sub = big[:, 1:4, 2:9, 1:7]   # planes themselves strided -> one contiguous copy
    np.random.seed(0)
    ref = tp.losses_composite(sub, lab[:, 1:4, 2:9, 1:7], True)
    np.random.seed(0)
    assert_losses_close(eco.losses_fn(sub, lab[:, 1:4, 2:9, 1:7], True), ref, what="strided planes")


def test_generic_organ_count_composite():
    """composite_set_theory with C != 3 (ratios supplied by the caller): composed from per-leaf kernel passes."""
    import ecologysemanticsegmentation_b200 as eco
    torch.manual_seed(4)
    for C, ratios in ((2, [1.0, 0.4]), (4, [1.0, 0.6, 0.3, 0.1])):
        p = (torch.rand(2, C, 8, 8) * 0.9 + 0.05).cuda()
        g = (torch.rand(2, C, 8, 8) > 0.5).float().cuda()
        np.random.seed(1)
        rl, rg = _ref("losses_composite", p, g, UP_ALL, True, 0, False, ratios)
        x = p.clone().requires_grad_(True)
        np.random.seed(1)
        ours = eco.losses_fn(x, g, True, relative_set_ratios=ratios)
        _combine(ours, UP_ALL).backward()
        assert_losses_close(ours, rl, what=f"C={C}")
        assert_grad_close(x.grad.cpu(), rg.cpu(), what=f"C={C}")


def test_intersection_and_union_loss():
    import ecologysemanticsegmentation_b200 as eco
    from oracle import torch_port as tp
    torch.manual_seed(6)
    sp, p = torch.rand(2, 1, 8, 8).cuda(), torch.rand(2, 1, 8, 8).cuda()
    g = (torch.rand(2, 1, 8, 8) > 0.5).float().cuda()
    assert_losses_close(eco.intersection_loss(sp, p, g), tp.leaf7(sp * p, g, 0, True))
    assert_losses_close(eco.union_loss(sp, p, g), tp.leaf7(g, tp.union_operand(sp, p), 0, True))
    assert isinstance(eco.union_loss(sp, p, g), eco.LossList)


def test_labels_that_require_grad_on_leaf_path():
    """Both arguments may carry grad."""
    import ecologysemanticsegmentation_b200 as eco
    from oracle import torch_port as tp
    torch.manual_seed(8)
    x0, g0 = torch.rand(2, 3, 8, 8) * 0.9 + 0.05, torch.rand(2, 3, 8, 8) * 0.9 + 0.05
    xr, gr = x0.clone().requires_grad_(True), g0.clone().requires_grad_(True)
    _combine(tp.losses_composite(xr, gr), UP_ALL).backward()
    x, g = x0.cuda().requires_grad_(True), g0.cuda().requires_grad_(True)
    _combine(eco.losses_fn(x, g), UP_ALL).backward()
    assert_grad_close(x.grad.cpu(), xr.grad, what="d/dx")
    assert_grad_close(g.grad.cpu(), gr.grad, what="d/dg")


def test_labels_that_require_grad_on_composite_path():
    """composite_set_theory=True with C == 3 and labels that require grad: the fused kernel only differentiates the
    predictions, so losses_fn composes the 21 leaves from pair-leaf passes (gradients w.r.t. both arguments)."""
    import ecologysemanticsegmentation_b200 as eco
    from oracle import torch_port as tp
    torch.manual_seed(18)
    x0, g0 = torch.rand(2, 3, 8, 8) * 0.9 + 0.05, torch.rand(2, 3, 8, 8) * 0.9 + 0.05
    xr, gr = x0.clone().requires_grad_(True), g0.clone().requires_grad_(True)
    np.random.seed(3)
    ref = tp.losses_composite(xr, gr, True)
    _combine(ref, UP_ALL).backward()
    x, g = x0.cuda().requires_grad_(True), g0.cuda().requires_grad_(True)
    np.random.seed(3)
    ours = eco.losses_fn(x, g, True)
    _combine(ours, UP_ALL).backward()
    assert_losses_close(ours, ref, what="composite, labels with grad")
    assert_grad_close(x.grad.cpu(), xr.grad, what="d/dx")
    assert_grad_close(g.grad.cpu(), gr.grad, what="d/dg")


def test_from_logits_composite_with_generic_organ_count():
    import ecologysemanticsegmentation_b200 as eco
    torch.manual_seed(14)
    ratios = [1.0, 0.6, 0.3, 0.1]
    z0 = torch.randn(2, 4, 8, 8)
    g = (torch.rand(2, 4, 8, 8) > 0.5).float()
    zr = z0.clone().requires_grad_(True)
    np.random.seed(1)
    from oracle import torch_port as tp
    ref = tp.losses_composite(torch.sigmoid(zr), g, True, 0, False, ratios)
    _combine(ref, UP_ALL).backward()
    z = z0.cuda().requires_grad_(True)
    np.random.seed(1)
    ours = eco.losses_fn(z, g.cuda(), True, relative_set_ratios=ratios, from_logits=True)
    _combine(ours, UP_ALL).backward()
    assert_losses_close(ours, ref, what="C=4 from logits")
    assert_grad_close(z.grad.cpu(), zr.grad, what="C=4 from logits")


def test_bf16_inputs_within_1e2():
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    z, g = make_inputs(4, 3, 64, 77)
    p16 = torch.sigmoid(z).to(torch.bfloat16)
    np.random.seed(0)
    rl, rg = _ref("losses_composite", p16.float(), g, UP, True)
    x = p16.cuda().requires_grad_(True)
    np.random.seed(0)
    ours = eco.losses_fn(x, g.cuda(), True)
    _combine(ours, UP).backward()
    assert x.grad.dtype == torch.bfloat16
    assert_losses_close(ours, rl, tol=TOL_BF16, what="bf16")
    assert_grad_close(x.grad.float().cpu(), rg, tol=TOL_BF16, what="bf16")
    # plain per-channel path with bf16 predictions
    rl, rg = _ref("losses_composite", p16.float(), g, UP, False)
    x = p16.cuda().requires_grad_(True)
    ours = eco.losses_fn(x, g.cuda(), False)
    _combine(ours, UP).backward()
    assert_losses_close(ours, rl, tol=TOL_BF16, what="bf16 plain")
    assert_grad_close(x.grad.float().cpu(), rg, tol=TOL_BF16, what="bf16 plain")


def test_unused_outputs_and_partial_backward():
    """train() weights some of the 7 outputs by 0 and .item()s the rest (train_multiclass.py:145-156)."""
    import ecologysemanticsegmentation_b200 as eco
    p = torch.rand(2, 3, 8, 8, device="cuda").requires_grad_(True)
    g = (torch.rand(2, 3, 8, 8, device="cuda") > 0.5).float()
    ce, bce, fl, dice, gdice, tw, fd = eco.losses_fn(p, g)
    assert float(ce) == 0.0
    (0 * fd + 1 * bce + 1 * (gdice + tw)).backward()
    assert p.grad is not None and torch.isfinite(p.grad).all()
    vals = [v.item() for v in (ce, bce, fl, dice, gdice, tw, fd)]
    assert all(np.isfinite(vals))


def test_empty_input_is_an_error():
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import _native
    with pytest.raises(_native.EcoLossError, match="empty"):
        eco.losses_fn(torch.zeros(0, 3, 4, 4, device="cuda"), torch.zeros(0, 3, 4, 4, device="cuda"))


def test_multi_threshold_beam_in_one_pass():
    """np.arange(0.8, 0.99, 0.01) (test_multiclass.py:64): 19 thresholds from one read == 19 single calls."""
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    torch.manual_seed(12)
    z = (torch.randn(4, 3, 64, 64) * 3).cuda()
    lab = (torch.rand(4, 3, 64, 64) > 0.5).float().cuda()
    thrs = np.arange(0.8, 0.99, step=0.01)
    assert len(thrs) == 19
    many, counts, _ = tmc.score_batch(z, lab, list(thrs), return_counts=True)
    assert many.shape == (19, 3)
    for k, t in enumerate(thrs):
        one, c1, _ = tmc.score_batch(z, lab, float(t), return_counts=True)
        assert torch.equal(c1[0], counts[k])
        assert torch.equal(one, many[k])


def test_reference_test_loop_semantics(tmp_path):
    """test(): mean over batches of per-batch Dice, returns None when the epoch directory exists."""
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    from oracle import torch_port as tp
    torch.manual_seed(2)
    C = len(tmc.ORGANS)
    batches = [(torch.rand(2, 3, 16, 16), (torch.rand(2, C, 16, 16) > 0.5).float(), ["a", "b"]) for _ in range(3)]
    net = torch.nn.Conv2d(3, C, 3, padding=1).cuda()
    out = tmc.test(net, batches, results_dir=str(tmp_path), saved_epoch=7)
    with torch.no_grad():
        ref = tp.eval_stream_dice([(net(x.cuda()), y.cuda()) for x, y, _ in batches])
    assert out.device.type == "cpu" and out.shape == (C,)
    assert_losses_close(out.numpy(), ref.numpy(), what="test() mean dice")
    assert tmc.test(net, batches, results_dir=str(tmp_path), saved_epoch=7) is None


def test_stats_are_additive_over_batch_shards_full_size():
    """Size-independent property at cfg2 size: sums of the two half-batches add up to the full-batch sums, and
    the gradient of shard k from the all-reduced sums equals rows k of the full-batch gradient."""
    from ecologysemanticsegmentation_b200 import ops
    from ecologysemanticsegmentation_b200.loss_composite import DEFAULT_RATIOS, composite3_leaf_scales, draw_pair_weights
    from ecologysemanticsegmentation_b200.synthetic import make_config
    z, g = make_config("cfg2")
    z, g = z.cuda(), g.cuda()
    full = ops.composite3_stats(z, g, True)
    a, b = ops.composite3_stats(z[:27], g[:27], True), ops.composite3_stats(z[27:], g[27:], True)
    # per-thread partials are fp32 (folded into fp64 every 8 iterations), so additivity holds to fp32-partial
    # rounding, far inside the 1e-5 budget; pixel counts and label counts are exact
    rel = ((a + b) - full).abs() / full.abs().clamp_min(1.0)
    assert float(rel.max()) < 1e-7, float(rel.max())
    assert torch.equal((a + b)[:7], full[:7])
    np.random.seed(0)
    scales = composite3_leaf_scales(draw_pair_weights(DEFAULT_RATIOS, False))
    up = torch.tensor(UP, dtype=torch.float32, device="cuda")
    l_full, jac_full, _ = ops.composite3_finalize(full, scales)
    l_sum, jac_sum, _ = ops.composite3_finalize(a + b, scales)
    assert torch.allclose(l_full, l_sum, rtol=1e-6)
    g_full = ops.composite3_grad(z, g, True, jac_full, up)
    g_b = ops.composite3_grad(z[27:], g[27:], True, jac_sum, up)
    assert_grad_close(g_b.cpu(), g_full[27:].cpu(), tol=1e-6, what="shard gradient from global sums")


def test_scoring_accepts_byte_and_bool_masks():
    """uint8 / bool masks: same counts and Dice as the float32 masks of the reference."""
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    torch.manual_seed(21)
    z = (torch.randn(3, 3, 64, 64) * 2).cuda()
    lab = (torch.rand(3, 3, 64, 64) > 0.55).cuda()
    d_f, c_f, s_f = tmc.score_batch(z, lab.float(), 0.8, return_counts=True)
    for masks in (lab, lab.to(torch.uint8)):
        d, c, s = tmc.score_batch(z, masks, 0.8, return_counts=True)
        assert torch.equal(c, c_f) and torch.equal(d, d_f)
        assert torch.allclose(s, s_f, rtol=1e-12)
    # odd plane size -> scalar path
    z2, l2 = z[:, :, :7, :9].contiguous(), lab[:, :, :7, :9].contiguous()
    assert torch.equal(tmc.score_batch(z2, l2, 0.8, return_counts=True)[1], tmc.score_batch(z2, l2.float(), 0.8, return_counts=True)[1])


def test_cfg4_sharding_property_on_one_gpu():
    """BASELINE configs[3]: 432 x 3 x 512 x 512 split into 8 shards of 54.  On one GPU the 8 shards are processed
    one after the other: the shard sums add up to the full-batch sums, and each shard's gradient from the global sums
    equals the matching rows of the single-launch full-batch gradient (the multi-process version of this, over NCCL /
    peer memory, is tests/test_gpu_distributed.py)."""
    from ecologysemanticsegmentation_b200 import fused, ops
    from ecologysemanticsegmentation_b200.distributed import shard_bounds
    from ecologysemanticsegmentation_b200.synthetic import make_config
    free, _ = torch.cuda.mem_get_info()
    if free < 8e9:
        pytest.skip("needs ~6 GB of device memory")
    z, g = make_config("cfg4")
    z, g = z.cuda(), g.cuda()
    np.random.seed(0)
    step = fused.CompositeLossStep(UP)
    losses_full, dz_full = step(z, g)
    acc = None
    for r in range(8):
        lo, hi = shard_bounds(432, 8, r)
        a = ops.composite3_stats(z[lo:hi], g[lo:hi], True)
        acc = a if acc is None else acc + a
    losses, jac, _ = ops.composite3_finalize(acc, step.scales)
    assert_losses_close(losses.cpu().numpy(), losses_full.cpu().numpy(), tol=1e-6, what="cfg4 sharded sums")
    for r in (0, 3, 7):
        lo, hi = shard_bounds(432, 8, r)
        dz = ops.composite3_grad(z[lo:hi], g[lo:hi], True, jac, step.upstream)
        assert_grad_close(dz.cpu(), dz_full[lo:hi].cpu(), tol=1e-6, what=f"cfg4 shard {r}")


def test_cfg5_frame_stream_scoring():
    """BASELINE configs[4]: a stream of 64 x 3 x 512 x 512 batches, sigmoid -> threshold -> per-class Dice per batch,
    mean over batches (test_multiclass.py:104), against the reference ops on the same device."""
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    from oracle import torch_port as tp
    batches = []
    for k in range(4):
        z, g = make_inputs(64, 3, 512, 105 + k)
        batches.append((z.cuda(), g.cuda()))
    for thr in (None, 0.8):
        ours = tmc.score_stream(batches, thr)
        ref = tp.eval_stream_dice(batches, thr)
        assert_losses_close(ours.cpu().numpy(), ref.numpy(), what=f"cfg5 stream thr={thr}")


@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (3, 2, 5, 7), (54, 3, 256, 256)])
def test_uint8_masks_bit_exact(shape):
    """test_multiclass.py:58,68-69,90-92: sigmoid -> optional threshold rule -> (t.numpy() * 255).astype(uint8).
    Bit-exact against the reference ops with the sigmoid evaluated on the same device."""
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    from oracle import counts as oc
    from oracle import torch_port as tp
    torch.manual_seed(31)
    z = (torch.randn(shape) * 3).cuda()
    lab = (torch.rand(shape) > 0.5).float().cuda()
    assert (tmc.to_uint8_masks(z).cpu().numpy() == oc.masks_u8(torch.sigmoid(z))).all()
    for thr in (0.8, 0.5, 0.93):
        ref = oc.masks_u8(tp.threshold_inplace(torch.sigmoid(z), thr))
        ours = tmc.to_uint8_masks(z, thr)
        assert ours.dtype == torch.uint8 and tuple(ours.shape) == shape
        assert (ours.cpu().numpy() == ref).all()
        assert set(np.unique(ref).tolist()) <= {0, 255}
    assert (tmc.to_uint8_masks(lab, inputs_are_probs=True).cpu().numpy() == oc.masks_u8(lab)).all()
    p = torch.rand(shape).cuda()
    assert (tmc.to_uint8_masks(p, inputs_are_probs=True).cpu().numpy() == oc.masks_u8(p)).all()
    # a channel slice (strided view) and bf16 probabilities
    if shape[1] > 1:
        assert (tmc.to_uint8_masks(z[:, 1:2], 0.8).cpu().numpy() == oc.masks_u8(tp.threshold_inplace(torch.sigmoid(z[:, 1:2]), 0.8))).all()
    pb = p.bfloat16()
    assert (tmc.to_uint8_masks(pb, inputs_are_probs=True).cpu().numpy() == oc.masks_u8(pb.float())).all()


def test_stream_scorer_equals_per_batch_scoring():
    """StreamScorer (one launch per batch into a slot, closed forms for all batches at the end) gives exactly the
    per-batch Dice of score_batch and the mean of score_stream -- thresholded and soft."""
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    from oracle import torch_port as tp
    torch.manual_seed(77)
    batches = [((torch.randn(4, 3, 32, 32) * 2).cuda(), (torch.rand(4, 3, 32, 32) > 0.6).float().cuda()) for _ in range(5)]
    for thr in (None, 0.8):
        sc = tmc.StreamScorer(3, 8, thr)
        for z, lab in batches:
            sc.add(z, lab)
        per = sc.per_batch()
        assert tuple(per.shape) == (5, 3)
        for k, (z, lab) in enumerate(batches):
            assert torch.equal(per[k], tmc.score_batch(z, lab, thr))
        ref = tp.eval_stream_dice([(z.cpu(), lab.cpu()) for z, lab in batches], thr)
        assert_losses_close(sc.result().cpu().numpy(), ref.numpy(), what=f"stream thr={thr}")
        sc.reset()
        sc.add(*batches[2])
        assert torch.equal(sc.result(), tmc.score_batch(*batches[2], thr))
    with pytest.raises(IndexError):
        sc = tmc.StreamScorer(3, 1, 0.8)
        sc.add(*batches[0])
        sc.add(*batches[1])


@pytest.mark.parametrize("case", [(54, 1, 256, 256, 0.0), (54, 1, 256, 256, 0.5), (3, 1, 17, 19, 0.5), (4, 2, 32, 32, 0.0),
                                  (54, 3, 256, 256, 0.0), (2, 5, 8, 24, 0.0)])
def test_leaf_loss_step_one_launch(case):
    """fused.LeafLossStep (eco_pair_fused: stats -> hand-over -> gradient in one cooperative launch) against the
    oracle on the same logits: cfg1 (C == 1, prediction in the gt slot, background_weight honoured) and plain C > 1."""
    from ecologysemanticsegmentation_b200 import fused
    from oracle import torch_port as tp
    n, c, h, w, bw = case
    torch.manual_seed(1000 + c + h)
    z0 = torch.randn(n, c, h, w)
    g = (torch.rand(n, c, h, w) > 0.5).float()
    for doubling, fn in ((2.0, tp.losses_composite), (1.0, tp.losses_train_multiclass)):
        zr = z0.clone().requires_grad_(True)
        ref = fn(torch.sigmoid(zr), g, False, bw)
        _combine(ref, UP_ALL).backward()
        step = fused.LeafLossStep(UP_ALL, doubling=doubling, background_weight=bw)
        for _ in range(3):   # several launches on the same workspace: generations / re-armed counters
            losses, grad = step(z0.cuda(), g.cuda())
        assert_losses_close(losses.cpu().numpy(), ref, what=f"leaf step {case} x{doubling}")
        assert_grad_close(grad.cpu(), zr.grad, what=f"leaf step {case} x{doubling}")


def test_leaf_loss_step_bf16_and_out_buffer():
    from ecologysemanticsegmentation_b200 import fused
    from oracle import torch_port as tp
    torch.manual_seed(5)
    z0 = torch.randn(6, 1, 64, 64)
    g = (torch.rand(6, 1, 64, 64) > 0.5).float()
    zr = z0.bfloat16().float().requires_grad_(True)
    ref = tp.losses_composite(torch.sigmoid(zr), g, False, 0.0)
    _combine(ref, UP_ALL).backward()
    out = torch.empty(6, 1, 64, 64, dtype=torch.bfloat16, device="cuda")
    losses, grad = fused.LeafLossStep(UP_ALL)(z0.bfloat16().cuda(), g.cuda(), out=out)
    assert grad.data_ptr() == out.data_ptr()
    assert_losses_close(losses.cpu().numpy(), ref, tol=TOL_BF16, what="leaf step bf16")
    assert_grad_close(grad.float().cpu(), zr.grad, tol=TOL_BF16, what="leaf step bf16")


@pytest.mark.parametrize("shape", [(3, 2, 7, 9), (4, 3, 64, 64), (8, 3, 256, 256)])
def test_threshold_beam_binning_kernel_edge_cases(shape):
    """5..20 thresholds go through the binning kernel: unsorted and duplicated thresholds, thresholds that hit
    probabilities exactly, byte masks, probabilities as input, odd plane sizes -- each row must equal the
    single-threshold call bit for bit (counts) and value for value (Dice)."""
    from ecologysemanticsegmentation_b200 import ops, test_multiclass as tmc
    torch.manual_seed(99)
    z = (torch.randn(shape) * 3).cuda()
    lab = (torch.rand(shape) > 0.5).float().cuda()
    prob = torch.sigmoid(z)
    exact_hits = [float(prob.flatten()[i]) for i in (0, 5, 17)]        # '>' is strict: these pixels must NOT count
    cases = [
        [0.9, 0.1, 0.5, 0.5, 0.97, 0.8],
        list(np.arange(0.8, 0.99, step=0.01)) + [0.5],                 # 20 thresholds
        exact_hits + [0.3, 0.6, 0.999999],
        [0.0, 1.0, -1.0, 2.0, 0.5],
    ]
    for thrs in cases:
        for labels in (lab, lab.to(torch.uint8)):
            many, counts, soft = tmc.score_batch(z, labels, thrs, return_counts=True)
            assert tuple(many.shape) == (len(thrs), shape[1])
            for k, t in enumerate(thrs):
                one, c1, s1 = tmc.score_batch(z, labels, float(t), return_counts=True)
                assert torch.equal(c1[0], counts[k]), (thrs, k)
                assert torch.equal(one, many[k])
            assert torch.allclose(soft, s1, rtol=1e-6)
        # probabilities as input: no sigmoid, plain strict compare
        many_p, counts_p, _ = tmc.score_batch(prob, lab, thrs, inputs_are_probs=True, return_counts=True)
        for k, t in enumerate(thrs):
            ref_out = (prob > torch.tensor(float(t), dtype=torch.float32)).long()
            for c in range(shape[1]):
                assert int(counts_p[k, c, 1]) == int(ref_out[:, c].sum())
                assert int(counts_p[k, c, 0]) == int((ref_out[:, c] * lab[:, c].long()).sum())
    thr_nan = torch.tensor([0.5, float("nan"), 0.7, 0.2, 0.9], device="cuda")
    c_nan, _ = ops.dice_counts(z, lab, thr_nan)
    assert int(c_nan[1, :, :2].abs().sum()) == 0                       # p > NaN is never true
