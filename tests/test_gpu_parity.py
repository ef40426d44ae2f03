"""Parity of the CUDA path against the oracle (= the reference, see oracle/__init__.py) on identical
seeded inputs.  Tolerances: loss scalars 1e-5 relative, gradients 1e-5 in max-norm and rel-L2 (fp32),
thresholded pixel counts bit-exact."""
import numpy as np
import pytest
import torch

from parity import assert_grad_close, assert_losses_close, TOL

pytestmark = pytest.mark.gpu

UP = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]          # bce + gdice + twersky + focal_dice (cfg2 combination)
UP_ALL = [0.3, 1.0, 0.7, 0.2, 1.0, 0.5, 1.0]      # every output carries gradient


def _inputs(n, c, s, seed, **kw):
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    return make_inputs(n, c, s, seed, **kw)


def _combine(losses, up):
    return sum(float(w) * l for w, l in zip(up, losses) if w != 0.0)


def _oracle(fn, x_cpu, g_cpu, up, *args, device="cpu", **kw):
    from oracle import torch_port as tp
    x = x_cpu.to(device).clone().requires_grad_(True)
    g = g_cpu.to(device)
    losses = getattr(tp, fn)(x, g, *args, **kw)
    _combine(losses, up).backward()
    return [float(v) for v in losses], x.grad.detach().cpu()


def _ours(x_cpu, g_cpu, up, *args, **kw):
    import ecologysemanticsegmentation_b200 as eco
    x = x_cpu.cuda().requires_grad_(True)
    g = g_cpu.cuda()
    losses = eco.losses_fn(x, g, *args, **kw)
    assert isinstance(losses, eco.LossList) and len(losses) == 7
    _combine(losses, up).backward()
    return [float(v) for v in losses], x.grad.detach().cpu()


@pytest.mark.parametrize("bw", [0, 0.5])
@pytest.mark.parametrize("shape", [(2, 1, 16, 16), (3, 1, 17, 19), (54, 1, 256, 256)])
def test_single_channel_leaf(shape, bw):
    z, g = _inputs(*shape[:2], shape[2], 101) if shape[2] == shape[3] else (torch.randn(shape), (torch.rand(shape) > 0.5).float())
    p = torch.sigmoid(z)
    ref_l, ref_g = _oracle("losses_composite", p, g, UP_ALL, False, bw)
    our_l, our_g = _ours(p, g, UP_ALL, False, bw)
    assert_losses_close(our_l, ref_l, what=f"leaf{shape} bw={bw}")
    assert_grad_close(our_g, ref_g, what=f"leaf{shape} bw={bw}")


@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (2, 2, 8, 24), (5, 4, 33, 7), (54, 3, 256, 256)])
def test_multichannel_plain(shape):
    torch.manual_seed(7)
    z = torch.randn(shape)
    g = (torch.rand(shape) > 0.5).float()
    p = torch.sigmoid(z)
    ref_l, ref_g = _oracle("losses_composite", p, g, UP_ALL, False, 0)
    our_l, our_g = _ours(p, g, UP_ALL, False, 0)
    assert_losses_close(our_l, ref_l, what=f"plain{shape}")
    assert_grad_close(our_g, ref_g, what=f"plain{shape}")


@pytest.mark.parametrize("up", [UP, UP_ALL])
@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (3, 3, 17, 19), (54, 3, 256, 256)])
def test_composite_on_probabilities(shape, up):
    if shape[2] == shape[3]:
        z, g = _inputs(shape[0], 3, shape[2], 102)
    else:
        torch.manual_seed(3)
        z, g = torch.randn(shape), (torch.rand(shape) > 0.5).float()
    p = torch.sigmoid(z)
    np.random.seed(0)
    ref_l, ref_g = _oracle("losses_composite", p, g, up, True)
    np.random.seed(0)
    our_l, our_g = _ours(p, g, up, True)
    assert_losses_close(our_l, ref_l, what=f"composite{shape}")
    assert_grad_close(our_g, ref_g, what=f"composite{shape}")


def test_composite_early_stopped_rng_stream():
    z, g = _inputs(4, 3, 32, 5)
    p = torch.sigmoid(z)
    np.random.seed(123)
    ref_l, ref_g = _oracle("losses_composite", p, g, UP, True, 0, True)
    state_ref = np.random.get_state()[1].copy()
    np.random.seed(123)
    our_l, our_g = _ours(p, g, UP, True, 0, True)
    state_ours = np.random.get_state()[1].copy()
    assert (state_ref == state_ours).all(), "numpy RNG stream consumed differently from the reference"
    assert_losses_close(our_l, ref_l)
    assert_grad_close(our_g, ref_g)


@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (54, 3, 256, 256)])
def test_composite_from_logits_same_device_oracle(shape):
    """from_logits: the oracle's sigmoid runs on the SAME device (torch CUDA) so that the |x_i-x_j| kink sees
    identical probability bits (SURVEY.md section 7 'abs() kink')."""
    import ecologysemanticsegmentation_b200 as eco
    from oracle import torch_port as tp
    z_cpu, g_cpu = _inputs(shape[0], 3, shape[2], 102)
    z = z_cpu.cuda().requires_grad_(True)
    np.random.seed(0)
    ref = tp.losses_composite(torch.sigmoid(z), g_cpu.cuda(), True)
    _combine(ref, UP).backward()
    ref_g = z.grad.detach().cpu()
    z2 = z_cpu.cuda().requires_grad_(True)
    np.random.seed(0)
    ours = eco.losses_fn(z2, g_cpu.cuda(), True, from_logits=True)
    _combine(ours, UP).backward()
    assert_losses_close(ours, ref, what="from_logits")
    assert_grad_close(z2.grad.cpu(), ref_g, what="from_logits")


def test_nonbinary_labels_take_exact_slow_path():
    torch.manual_seed(11)
    p = torch.rand(2, 3, 16, 16) * 0.98 + 0.01
    g = (torch.rand(2, 3, 16, 16) > 0.5).float()
    g[0, 1, 3, 4] = 0.25
    g[1, 2, 0, 0] = 0.6
    g[1, 0, 5, 5] = 0.5
    np.random.seed(0)
    ref_l, ref_g = _oracle("losses_composite", p, g, UP_ALL, True)
    np.random.seed(0)
    our_l, our_g = _ours(p, g, UP_ALL, True)
    assert_losses_close(our_l, ref_l, what="nonbinary labels")
    assert_grad_close(our_g, ref_g, what="nonbinary labels")


@pytest.mark.parametrize("thr", [None, 0.8, 0.9])
@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (3, 2, 17, 19), (8, 3, 256, 256)])
def test_eval_counts_bit_exact(shape, thr):
    from ecologysemanticsegmentation_b200 import ops
    from oracle import counts as oc
    torch.manual_seed(5)
    z = torch.randn(shape) * 2
    lab = (torch.rand(shape) > 0.6).float()
    zc, lc = z.cuda(), lab.cuda()
    thr_t = None if thr is None else torch.tensor([thr], dtype=torch.float32, device="cuda")
    counts, soft = ops.dice_counts(zc, lc, thr_t)
    soft_ref = oc.batch_soft_sums(zc, lc)
    np.testing.assert_allclose(soft.cpu().numpy(), soft_ref, rtol=1e-6)
    if thr is not None:
        ref = oc.batch_counts(zc, lc, thr)  # same-device sigmoid bits
        assert (counts[0].cpu().numpy() == ref).all(), (counts[0].cpu().numpy(), ref)


@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (3, 3, 17, 19), (54, 3, 256, 256)])
def test_fused_single_launch_matches_split_path(shape):
    """The cooperative single-launch kernel must agree with the stats/finalize/grad path bit for bit on the
    gradient coefficients' inputs (same sums, same closed forms) and to 1e-6 on outputs."""
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    torch.manual_seed(3)
    z = torch.randn(shape).cuda()
    g = (torch.rand(shape) > 0.5).float().cuda()
    np.random.seed(0)
    step = CompositeLossStep(UP)
    losses, dz = step(z, g)
    z2 = z.clone().requires_grad_(True)
    np.random.seed(0)
    ref = eco.losses_fn(z2, g, True, from_logits=True)
    _combine(ref, UP).backward()
    assert_losses_close(losses, ref, tol=1e-6, what="fused vs split")
    assert_grad_close(dz.cpu(), z2.grad.cpu(), tol=1e-6, what="fused vs split")


@pytest.mark.parametrize("nthr", [1, 3, 19])
def test_thresholded_dice_with_fractional_labels_matches_reference(nthr):
    """Masks resized by the dataset are not 0/1 (ess/dataset/fish/fish_suim.py:60-74, fish_deepfish_segment.py:71-84).  The
    reference's thresholded Dice uses the real label values, 2 sum(out*lab) / (sum out + sum lab^2)
    (ess/test_multiclass.py:68-69,80 -> ess/loss_functions.py:55-57); so must the scoring here -- class 1 has fractional
    labels (the exact second pass), classes 0 and 2 stay binary (integer counts), in one call."""
    from ecologysemanticsegmentation_b200 import ops, test_multiclass as tmc
    from oracle import torch_port as tp
    torch.manual_seed(21)
    z = (torch.randn(3, 3, 40, 52) * 2.5).cuda()
    lab = (torch.rand(3, 3, 40, 52) > 0.5).float()
    frac = torch.rand(3, 40, 52)
    lab[:, 1] = torch.where(frac < 0.2, torch.round(frac * 5 * 255) / 255, lab[:, 1])   # k/255 values among the 0/1
    lab[0, 1, 0, 0] = -1.0 / 255                                                       # the datasets' "missing organ" value
    lab = lab.cuda()
    thrs = [0.8] if nthr == 1 else list(np.arange(0.8, 0.99, 0.01)[:nthr])
    got = tmc.score_batch(z, lab, thrs if nthr > 1 else thrs[0])
    got = got.reshape(nthr, 3).cpu().numpy()
    for k, t in enumerate(thrs):
        ref = [float(v) for v in tp.eval_batch_dice(z, lab, float(np.float32(t)))]
        assert_losses_close(got[k], ref, what=f"fractional labels, T={t:.2f}")
    # the plain (integer-only) entry point marks the class instead of returning a silently different number
    counts, _ = ops.dice_counts(z, lab, torch.tensor(thrs, dtype=torch.float32, device="cuda"))
    c = counts.cpu().numpy()
    assert (c[:, 1, 2] == -1).all() and (c[:, 0, 2] >= 0).all() and (c[:, 2, 2] >= 0).all()


# ----------------------------------------------------------------------------------------------
# the sequential model's test: sigmoid -> prediction un-union -> soft Dice in ONE read
# (ess/test_multiclass_sequential_densenetloss.py:62,66,97-99)
# ----------------------------------------------------------------------------------------------
def _ununion_reference(zc, lc, probs=False):
    """The reference's own sequence on the same device: F.sigmoid, the in-place un-union, then dice_loss's sums per class
    (float64 here so that the comparison is about the per-element values, not the summation order)."""
    from oracle import torch_port as tp
    out = zc.float().clone() if probs else torch.sigmoid(zc.float())
    out = tp.union_sets_descending(out, reverse=True).double()
    lab = lc.double()
    C = lab.shape[1]
    sums = np.array([[float((out[:, c] * lab[:, c]).sum()), float(out[:, c].sum()), float((lab[:, c] ** 2).sum())] for c in range(C)])
    dice = torch.stack([-tp.pair_dice(out[:, c:c + 1].float(), lc[:, c:c + 1].float(), background_weight=0) for c in range(C)])
    return sums, dice


@pytest.mark.parametrize("labels_dtype", [torch.float32, torch.uint8])
@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (3, 3, 17, 19), (2, 4, 32, 32), (2, 5, 9, 7), (3, 2, 16, 16), (2, 1, 16, 16),
                                   (8, 3, 256, 256)])
def test_sequential_test_ununion_fused_into_soft_dice(shape, labels_dtype):
    from ecologysemanticsegmentation_b200 import ops, test_multiclass_sequential_densenetloss as seq
    torch.manual_seed(11)
    z = (torch.randn(shape) * 2).cuda()
    lab = (torch.rand(shape) > 0.6).to(labels_dtype).cuda()
    keep = z.clone()
    _, soft, _ = ops.dice_counts_ex(z, lab, None, ununion_preds=True)
    assert torch.equal(z, keep)                                     # the predictions are not rewritten
    ref_sums, ref_dice = _ununion_reference(z, lab)
    np.testing.assert_allclose(soft.cpu().numpy(), ref_sums, rtol=2e-6, atol=1e-9)
    d = seq.score_batch(z, lab)
    np.testing.assert_allclose(d.cpu().numpy(), ref_dice.cpu().numpy(), rtol=1e-5)
    # ... and it is the same thing as un-unioning in place first (the stand-alone kernel) and scoring the probabilities
    from ecologysemanticsegmentation_b200 import subsets_union
    pr = subsets_union.return_union_sets_descending_order(torch.sigmoid(z), reverse=True)
    _, soft2, _ = ops.dice_counts_ex(pr, lab, None, inputs_are_probs=True)
    np.testing.assert_allclose(soft.cpu().numpy(), soft2.cpu().numpy(), rtol=2e-6, atol=1e-9)


def test_sequential_test_ununion_on_probabilities_and_bf16():
    from ecologysemanticsegmentation_b200 import ops
    torch.manual_seed(12)
    p = torch.rand(3, 4, 24, 24).cuda()
    lab = (torch.rand(3, 4, 24, 24) > 0.5).float().cuda()
    _, soft, _ = ops.dice_counts_ex(p, lab, None, inputs_are_probs=True, ununion_preds=True)
    ref_sums, _ = _ununion_reference(p, lab, probs=True)
    np.testing.assert_allclose(soft.cpu().numpy(), ref_sums, rtol=2e-6, atol=1e-9)
    zb = (torch.randn(2, 3, 32, 32) * 2).cuda().bfloat16()
    labb = (torch.rand(2, 3, 32, 32) > 0.5).float().cuda()
    _, softb, _ = ops.dice_counts_ex(zb, labb, None, ununion_preds=True)
    ref_b, _ = _ununion_reference(zb, labb)                          # the oracle on the widened bf16 logits
    np.testing.assert_allclose(softb.cpu().numpy(), ref_b, rtol=2e-6, atol=1e-9)


def test_sequential_test_ununion_refuses_thresholds():
    from ecologysemanticsegmentation_b200 import ops, test_multiclass
    from ecologysemanticsegmentation_b200._native import EcoLossError
    z = torch.randn(2, 3, 16, 16).cuda()
    lab = (torch.rand(2, 3, 16, 16) > 0.5).float().cuda()
    thr = torch.tensor([0.8], dtype=torch.float32, device="cuda")
    with pytest.raises(EcoLossError):
        ops.dice_counts_ex(z, lab, thr, ununion_preds=True)
    with pytest.raises(ValueError):
        test_multiclass.score_batch(z, lab, 0.8, ununion=True)


def test_sequential_test_function_matches_reference_loop(tmp_path):
    """test() of the sequential module against the reference's batch loop (:50-99, :137-140 mean of per-batch Dice)."""
    from ecologysemanticsegmentation_b200 import test_multiclass_sequential_densenetloss as seq
    from oracle import torch_port as tp
    torch.manual_seed(13)
    batches = [(torch.randn(2, 3, 32, 32), (torch.rand(2, 3, 32, 32) > 0.5).float(), [0, 1]) for _ in range(3)]

    class Net(torch.nn.Module):
        def forward(self, x):
            return x * 1.5
    got = seq.test(Net(), batches, results_dir=str(tmp_path / "r"), saved_epoch=3)
    acc = torch.zeros(3, dtype=torch.float64)
    for x, lab, _ in batches:
        out = tp.union_sets_descending(torch.sigmoid(x.cuda() * 1.5), reverse=True)
        acc += torch.stack([-tp.pair_dice(out[:, c:c + 1], lab.cuda()[:, c:c + 1], background_weight=0) for c in range(3)]).double().cpu()
    np.testing.assert_allclose(got.numpy(), (acc / 3).float().numpy(), rtol=1e-5)
    assert seq.test(Net(), batches, results_dir=str(tmp_path / "r"), saved_epoch=3) is None   # "Test already done"


# ----------------------------------------------------------------------------------------------
# soft-label cross entropy over the channel dim (loss_functions.py:26-30,44), every kernel variant: 128-bit kernels for 2..4
# channels on aligned planes, the generic kernel otherwise (other channel counts, H*W % 4 != 0, strided slices)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bw", [0, 0.3])
@pytest.mark.parametrize("shape", [(2, 2, 16, 16), (3, 3, 32, 40), (2, 4, 8, 24), (2, 5, 16, 16), (3, 3, 17, 19), (2, 7, 5, 3),
                                   (54, 3, 256, 256)])
def test_soft_label_cross_entropy_value_and_both_gradients(shape, bw):
    import ecologysemanticsegmentation_b200 as eco
    from oracle import torch_port as tp
    torch.manual_seed(90 + shape[1])
    a0 = torch.rand(shape).cuda()
    b0 = (torch.randn(shape) * 2).cuda()
    ar, br = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    ref = tp.pair_soft_ce(ar, br, bw)
    ref.backward()
    ao, bo = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    got = eco.loss_functions.cross_entropy_loss(ao, bo, background_weight=bw)
    got.backward()
    assert abs(float(got) - float(ref)) <= TOL * abs(float(ref))
    assert_grad_close(ao.grad.cpu(), ar.grad.cpu(), what=f"soft CE d/dgt {shape} bw={bw}")
    assert_grad_close(bo.grad.cpu(), br.grad.cpu(), what=f"soft CE d/dpred {shape} bw={bw}")


def test_soft_label_cross_entropy_slices_and_bf16():
    import ecologysemanticsegmentation_b200 as eco
    from oracle import torch_port as tp
    torch.manual_seed(91)
    a0 = torch.rand(3, 6, 16, 24).cuda()
    b0 = (torch.randn(3, 6, 16, 24) * 2).cuda()
    # channel slices of a larger tensor: strided planes, three channels -> the 128-bit kernels on non-contiguous input
    ar, br = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    tp.pair_soft_ce(ar[:, 1:4], br[:, 1:4], 0.3).backward()
    ao, bo = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    eco.loss_functions.cross_entropy_loss(ao[:, 1:4], bo[:, 1:4], background_weight=0.3).backward()
    assert_grad_close(bo.grad.cpu(), br.grad.cpu(), what="soft CE on channel slices")
    assert_grad_close(ao.grad.cpu(), ar.grad.cpu(), what="soft CE on channel slices (gt)")
    # bf16 predictions
    bb = b0[:, :3].bfloat16()
    brf = bb.float().requires_grad_(True)
    ref = tp.pair_soft_ce(a0[:, :3], brf, 0.0)
    ref.backward()
    bq = bb.clone().requires_grad_(True)
    got = eco.loss_functions.cross_entropy_loss(a0[:, :3].contiguous(), bq, background_weight=0.0)
    got.backward()
    assert abs(float(got) - float(ref)) <= 1e-2 * abs(float(ref))
    assert_grad_close(bq.grad.float().cpu(), brf.grad.cpu(), tol=1e-2, what="soft CE bf16")
