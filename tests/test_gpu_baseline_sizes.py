"""Parity at BASELINE.json's OWN sizes against the same-device oracle (reference semantics in eager torch ops).

cfg3 (54x3x1024x1024): 56.6 M pixels per class -- the counts exceed 2**24, the regime where the reference's own fp32
``torch.sum`` is inexact and where the packed 16+16-bit counters of the scoring kernels fold;
cfg4's per-GPU shard (54x3x512x512) for the fused composite step; cfg2's size in bf16.
Reference lines: ess/test_multiclass.py:58,64,68-69,80-82; ess/loss_composite.py:21-94; ess/train_multiclass.py:134-147.
"""
import numpy as np
import pytest
import torch

from parity import assert_grad_close, assert_losses_close, TOL, TOL_BF16

pytestmark = pytest.mark.gpu

UP_CFG2 = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]


@pytest.fixture(scope="module")
def cfg3():
    from ecologysemanticsegmentation_b200.synthetic import make_config
    z, g = make_config("cfg3")
    zc, gc = z.cuda(), g.cuda()
    del z, g
    yield zc, gc
    del zc, gc
    torch.cuda.empty_cache()


def test_cfg3_full_size_counts_bit_exact(cfg3):
    """sigmoid -> strict '>' T -> per-class (I, |out|, |lab|), integer-equal to the reference's thresholding ops on this
    device summed in int64; soft sums within 1e-6; Dice values within 1e-5."""
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    from oracle import counts as oc
    zc, gc = cfg3
    assert zc.shape == (54, 3, 1024, 1024)
    for thr in (0.8, 0.9):
        d, counts, soft = tmc.score_batch(zc, gc, thr, return_counts=True)
        ref = oc.batch_counts(zc, gc, thr)
        assert ref[:, 2].min() > 2 ** 22 and ref[0, 2] > 2 ** 24, "the case must exceed the fp32-exact range"
        assert (counts[0].cpu().numpy() == ref).all(), (thr, counts[0].cpu().numpy(), ref)
        assert_losses_close(d.cpu().numpy(), oc.dice_from_counts(ref), tol=TOL, what=f"cfg3 dice@{thr}")
        np.testing.assert_allclose(soft.cpu().numpy(), oc.batch_soft_sums(zc, gc), rtol=1e-6)
    # the live (un-thresholded) path of test(): soft Dice vs the reference ops in float64
    sd = tmc.score_batch(zc, gc)
    s = oc.batch_soft_sums(zc, gc)
    assert_losses_close(sd.cpu().numpy(), (2 * s[:, 0] + 1e-7) / (s[:, 1] + s[:, 2] + 1e-7), tol=TOL, what="cfg3 soft dice")


def test_cfg3_full_size_beam_19_thresholds(cfg3):
    """np.arange(0.8, 0.99, 0.01) (ess/test_multiclass.py:64) in ONE read at the full size: every threshold's counts
    equal the oracle's (same float32 threshold value, same device)."""
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    from oracle import counts as oc
    zc, gc = cfg3
    thrs = np.arange(0.8, 0.99, step=0.01)
    assert len(thrs) == 19
    many, counts, _ = tmc.score_batch(zc, gc, list(thrs), return_counts=True)
    got = counts.cpu().numpy()
    assert got.shape == (19, 3, 3)
    for k in (0, 1, 5, 9, 13, 17, 18):
        ref = oc.batch_counts(zc, gc, float(np.float32(thrs[k])))
        assert (got[k] == ref).all(), (k, thrs[k], got[k], ref)
        assert_losses_close(many[k].cpu().numpy(), oc.dice_from_counts(ref), tol=TOL, what=f"beam dice@{thrs[k]:.2f}")
    # monotone in the threshold: |out| can only shrink
    assert (np.diff(got[:, :, 1], axis=0) <= 0).all()


def test_cfg3_full_size_byte_masks_equal_float_masks(cfg3):
    """uint8 {0,1} masks must give the same counts as their float32 form."""
    from ecologysemanticsegmentation_b200 import ops
    zc, gc = cfg3
    thr = torch.tensor([0.8], dtype=torch.float32, device="cuda")
    c32, s32 = ops.dice_counts(zc, gc, thr)
    c8, s8 = ops.dice_counts(zc, gc.to(torch.uint8), thr)
    assert torch.equal(c32, c8)
    np.testing.assert_allclose(s8.cpu().numpy(), s32.cpu().numpy(), rtol=1e-12)


def _oracle_composite(z, g, up):
    from oracle import torch_port as tp
    zr = z.clone().requires_grad_(True)
    np.random.seed(0)
    ref = tp.losses_composite(torch.sigmoid(zr), g, True)
    sum(w * l for w, l in zip(up, ref) if w).backward()
    return [float(v) for v in ref], zr.grad


def test_cfg4_shard_fused_step_vs_oracle():
    """BASELINE configs[3]'s per-GPU shard, 54x3x512x512: the one-launch fused step against the oracle's autograd on the
    same device (identical sigmoid bits), 1e-5 on the 7 losses and on the gradient (max-norm and rel-L2)."""
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from ecologysemanticsegmentation_b200.synthetic import make_config
    z, g = make_config("cfg4", n=54)
    zc, gc = z.cuda(), g.cuda()
    np.random.seed(0)
    losses, dz = CompositeLossStep(UP_CFG2)(zc, gc)
    rl, rg = _oracle_composite(zc, gc, UP_CFG2)
    assert_losses_close(losses.cpu().numpy(), rl, tol=TOL, what="cfg4 shard fused")
    d = (dz - rg).double()
    mx = float(d.abs().max() / rg.abs().max())
    l2 = float(d.norm() / rg.double().norm())
    assert mx <= TOL and l2 <= TOL, (mx, l2)
    del rg, dz
    torch.cuda.empty_cache()


def test_cfg2_full_size_bf16_within_1e2():
    """north_star: 1e-2 in bf16.  bf16 logits (and bf16 gradient out) at cfg2's full size against the oracle evaluated in
    fp32 on the same (bf16-rounded) logits."""
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from ecologysemanticsegmentation_b200.synthetic import make_config
    z, g = make_config("cfg2")
    z16 = z.to(torch.bfloat16).cuda()
    gc = g.cuda()
    np.random.seed(0)
    losses, dz = CompositeLossStep(UP_CFG2)(z16, gc)
    assert dz.dtype == torch.bfloat16
    rl, rg = _oracle_composite(z16.float(), gc, UP_CFG2)
    assert_losses_close(losses.cpu().numpy(), rl, tol=TOL_BF16, what="cfg2 bf16 fused")
    d = (dz.float() - rg).double()
    assert float(d.abs().max() / rg.abs().max()) <= TOL_BF16
    assert float(d.norm() / rg.double().norm()) <= TOL_BF16
    torch.cuda.empty_cache()


def test_cfg2_full_size_fused_step_vs_oracle():
    """The headline workload itself (54x3x256x256 fp32, from logits, one launch) against the same-device oracle."""
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    from ecologysemanticsegmentation_b200.synthetic import make_config
    z, g = make_config("cfg2")
    zc, gc = z.cuda(), g.cuda()
    np.random.seed(0)
    losses, dz = CompositeLossStep(UP_CFG2)(zc, gc)
    rl, rg = _oracle_composite(zc, gc, UP_CFG2)
    assert float(losses[0]) == 0.0
    assert_losses_close(losses.cpu().numpy(), rl, tol=TOL, what="cfg2 fused")
    assert_grad_close(dz.cpu(), rg.cpu(), tol=TOL, what="cfg2 fused")
