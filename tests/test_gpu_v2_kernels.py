"""The second-generation composite kernels (csrc/eco_composite_v2.cuh: TMA tile pipeline, scalar pass 1, pass 2 with
the linear BCE/focal sums, integer grid sums) against the oracle and against the first-generation split path.

They serve fp32 logits with 16-byte aligned planes (H*W % 4 == 0); every shape here is chosen to hit one of their
edge cases: short last tiles (H*W % 1024 != 0), fewer tiles than SMs, planes smaller than one tile, ties of the
|x_i - x_j| kink, saturated logits, labels other than 0/1, arbitrary (also negative) leaf scales, workspace re-arming.
Tolerances as everywhere: losses 1e-5 relative, gradients 1e-5 in max-norm and rel-L2."""
import numpy as np
import pytest
import torch

from parity import assert_grad_close, assert_losses_close

pytestmark = pytest.mark.gpu

UP = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]          # bce + gdice + twersky + focal_dice (cfg2 combination)
UP_ALL = [0.3, 1.0, 0.7, 0.2, 1.0, 0.5, 1.0]      # every output carries gradient (focal too)
UP_DICE = [0.0, 0.0, 0.0, 1.0, 1.0, 0.0, 1.0]     # no BCE / focal gradient: the leanest pass-2 variant
UP_FOCAL = [0.0, 0.0, 1.0, 0.0, 0.0, 1.0, 0.0]    # focal gradient without the BCE one


def _combine(losses, up):
    return sum(float(w) * l for w, l in zip(up, losses) if w != 0.0)


def _nested_masks(shape, gen):
    n, _, h, w = shape
    u = torch.rand((n, 1, h, w), generator=gen)
    return torch.cat([(u < 0.5).float(), (u < 0.5 * 0.43197708).float(), (u < 0.5 * 0.22319692).float()], dim=1)


def _oracle_same_device(z_cpu, g_cpu, up):
    """The reference's ops on the same GPU (identical sigmoid bits for the |.| kink), fp32 autograd."""
    from oracle import torch_port as tp
    z = z_cpu.cuda().requires_grad_(True)
    np.random.seed(0)
    ref = tp.losses_composite(torch.sigmoid(z), g_cpu.cuda(), True)
    _combine(ref, up).backward()
    return [float(v) for v in ref], z.grad.detach().cpu()


def _fused(z_cpu, g_cpu, up):
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    np.random.seed(0)
    step = CompositeLossStep(up)
    losses, dz = step(z_cpu.cuda(), g_cpu.cuda())
    return [float(v) for v in losses.cpu()], dz.cpu()


@pytest.mark.parametrize("up", [UP, UP_ALL, UP_DICE, UP_FOCAL])
@pytest.mark.parametrize("shape", [
    (2, 3, 36, 28),      # H*W = 1008: every plane is one short tile
    (3, 3, 100, 100),    # 9 full tiles + a 784-pixel tail per plane
    (1, 3, 64, 64),      # 4 tiles: far fewer than SMs
    (200, 3, 32, 32),    # 200 planes of exactly one tile
    (7, 3, 128, 136),    # 17 tiles per plane, the CTA ranges straddle images
])
def test_fused_v2_vs_oracle(shape, up):
    gen = torch.Generator().manual_seed(shape[0] * 1000 + shape[2])
    z = torch.randn(shape, generator=gen)
    g = _nested_masks(shape, gen)
    ref_l, ref_g = _oracle_same_device(z, g, up)
    our_l, our_g = _fused(z, g, up)
    assert_losses_close(our_l, ref_l, what=f"fused v2 {shape}")
    assert_grad_close(our_g, ref_g, what=f"fused v2 {shape}")


def test_fused_v2_ties_and_saturation():
    """Exact and near ties of the probabilities (sign(0) = 0 of torch.abs' backward, sign from ATen's sigmoid bits)
    and logits deep in the saturated range."""
    gen = torch.Generator().manual_seed(77)
    shape = (4, 3, 64, 64)
    z = torch.randn(shape, generator=gen) * 4.0
    g = (torch.rand(shape, generator=gen) > 0.5).float()
    z[0, 1, :8] = z[0, 0, :8]                      # exact ties of channels 0 and 1
    z[1, 2, 8:16] = z[1, 1, 8:16]                  # exact ties of channels 1 and 2
    z[2, 2, :4] = z[2, 0, :4] + 1e-7               # near ties: one-ulp differences after the sigmoid
    z[3, :, :2] = 18.0                             # all three saturated high (p == 1 in fp32)
    z[3, :, 2:4] = -30.0                           # all three saturated low
    ref_l, ref_g = _oracle_same_device(z, g, UP_ALL)
    our_l, our_g = _fused(z, g, UP_ALL)
    assert_losses_close(our_l, ref_l, what="ties/saturation")
    assert_grad_close(our_g, ref_g, what="ties/saturation")


def test_fused_v2_nonbinary_labels():
    gen = torch.Generator().manual_seed(5)
    shape = (3, 3, 48, 64)
    z = torch.randn(shape, generator=gen)
    g = (torch.rand(shape, generator=gen) > 0.5).float()
    g[0, 1, 3, 4] = 0.25
    g[1, 2, 0, 0] = 0.6
    g[2, 0, 5, 5] = 0.5
    g[2, 1, 40:, 60:] = 0.125
    ref_l, ref_g = _oracle_same_device(z, g, UP_ALL)
    our_l, our_g = _fused(z, g, UP_ALL)
    assert_losses_close(our_l, ref_l, what="non-binary labels")
    assert_grad_close(our_g, ref_g, what="non-binary labels")


def test_fused_v2_matches_split_path_with_arbitrary_scales():
    """Arbitrary leaf scales, some negative, through the C ABI: the fused v2 kernel (BCE / focal sums weighted in
    pass 2) must agree with statistics -> finalize -> gradient (first-generation statistics kernel)."""
    from ecologysemanticsegmentation_b200 import ops
    gen = torch.Generator().manual_seed(9)
    shape = (5, 3, 72, 96)
    z = torch.randn(shape, generator=gen).cuda()
    g = (torch.rand(shape, generator=gen) > 0.5).float().cuda()
    scales = torch.tensor([2.0, 1.5, -0.75] + [0.3 * (k + 1) * (-1.0 if k % 5 == 3 else 1.0) for k in range(18)],
                          dtype=torch.float64, device="cuda")
    up = torch.tensor(UP_ALL, dtype=torch.float32, device="cuda")
    losses_f, grad_f = ops.composite3_fused(z, g, scales, up, True)
    acc = ops.composite3_stats(z, g, True)
    losses_s, jac, _ = ops.composite3_finalize(acc, scales)
    grad_s = ops.composite3_grad(z, g, True, jac, up)
    assert_losses_close(losses_f, losses_s, tol=2e-6, what="fused v2 vs split, arbitrary scales")
    assert_grad_close(grad_f.cpu(), grad_s.cpu(), tol=2e-6, what="fused v2 vs split, arbitrary scales")


def test_fused_v2_is_deterministic_and_rearms_its_workspace():
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    gen = torch.Generator().manual_seed(21)
    shape = (54, 3, 128, 128)
    z = torch.randn(shape, generator=gen).cuda()
    g = _nested_masks(shape, gen).cuda()
    np.random.seed(0)
    step = CompositeLossStep(UP)
    first_l, first_g = step(z, g)
    first_l, first_g = first_l.clone(), first_g.clone()
    for _ in range(5):
        l, dz = step(z, g)
        assert torch.equal(l, first_l), "loss values changed between identical steps"
        assert torch.equal(dz, first_g), "gradient changed between identical steps"
    # a different batch in between must not leave anything behind in the integer accumulators
    z2 = torch.randn(shape, generator=gen).cuda()
    step(z2, g)
    l, dz = step(z, g)
    assert torch.equal(l, first_l) and torch.equal(dz, first_g)


def test_grad_v2_all_upstream_variants_match_first_generation():
    """The stand-alone pass 2 (autograd path) against the first-generation scalar kernel.  The gradient of a pixel
    depends on that pixel and on the global coefficients only, so the same coefficients (jac) applied to a crop
    whose planes are NOT 16-byte aligned (H*W % 4 != 0 -> scalar kernels) must give the same values."""
    from ecologysemanticsegmentation_b200 import ops
    gen = torch.Generator().manual_seed(33)
    n, h, w = 3, 40, 52
    z = torch.randn((n, 3, h, w), generator=gen).cuda()
    g = (torch.rand((n, 3, h, w), generator=gen) > 0.5).float().cuda()
    z[0, 1, :4] = z[0, 0, :4]   # ties
    scales = torch.tensor([2.0] * 3 + [2.0 * (1 + 0.1 * k) for k in range(18)], dtype=torch.float64, device="cuda")
    acc = ops.composite3_stats(z, g, True)
    _, jac, _ = ops.composite3_finalize(acc, scales)
    zc, gc = z[:, :, :39, :51].contiguous(), g[:, :, :39, :51].contiguous()
    assert (39 * 51) % 4 != 0
    for up_list in (UP, UP_ALL, UP_DICE, UP_FOCAL):
        up = torch.tensor(up_list, dtype=torch.float32, device="cuda")
        ours = ops.composite3_grad(z, g, True, jac, up)
        ref = ops.composite3_grad(zc, gc, True, jac, up)
        assert_grad_close(ours[:, :, :39, :51].cpu(), ref.cpu(), tol=2e-6, what=f"grad v2 vs scalar, up={up_list}")


# ---- the plain multi-class loss step (train_multiclass.losses_fn for C == 3) in one launch ---------------------------
@pytest.mark.parametrize("doubling", [1.0, 2.0])
@pytest.mark.parametrize("up", [UP, UP_ALL, UP_DICE, UP_FOCAL])
@pytest.mark.parametrize("shape", [(2, 3, 36, 28), (3, 3, 100, 100), (54, 3, 64, 64), (5, 3, 128, 136)])
def test_multiclass_step_vs_oracle(shape, up, doubling):
    """fused.MulticlassLossStep against the reference's ops on the same GPU: train_multiclass.losses_fn (doubling 1,
    train_multiclass.py:253-274) and loss_composite.losses_fn(composite_set_theory=False) (doubling 2, :28-40)."""
    from ecologysemanticsegmentation_b200.fused import MulticlassLossStep
    from oracle import torch_port as tp
    gen = torch.Generator().manual_seed(shape[0] * 31 + shape[3])
    z_cpu = torch.randn(shape, generator=gen) * 1.5
    g_cpu = _nested_masks(shape, gen)
    z = z_cpu.cuda().requires_grad_(True)
    fn = tp.losses_train_multiclass if doubling == 1.0 else tp.losses_composite
    ref = fn(torch.sigmoid(z), g_cpu.cuda(), False)
    _combine(ref, up).backward()
    losses, dz = MulticlassLossStep(up, doubling=doubling)(z_cpu.cuda(), g_cpu.cuda())
    assert_losses_close(losses.cpu(), [float(v) for v in ref], what=f"multiclass step {shape}")
    assert_grad_close(dz.cpu(), z.grad.cpu(), what=f"multiclass step {shape}")


def test_multiclass_step_matches_drop_in_and_is_deterministic():
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import train_multiclass as tm
    from ecologysemanticsegmentation_b200.fused import MulticlassLossStep
    gen = torch.Generator().manual_seed(4)
    shape = (54, 3, 128, 128)
    z = torch.randn(shape, generator=gen).cuda()
    g = (torch.rand(shape, generator=gen) > 0.6).float().cuda()
    step = MulticlassLossStep(UP_ALL)
    l0, d0 = step(z, g)
    l0, d0 = l0.clone(), d0.clone()
    for _ in range(3):
        l, d = step(z, g)
        assert torch.equal(l, l0) and torch.equal(d, d0)
    z2 = z.clone().requires_grad_(True)
    ref = tm.losses_fn(z2, g, from_logits=True)     # the drop-in callable: statistics -> finalize -> gradient kernels
    _combine(ref, UP_ALL).backward()
    assert_losses_close(l0.cpu(), [float(v) for v in ref], tol=2e-6, what="multiclass step vs drop-in")
    assert_grad_close(d0.cpu(), z2.grad.cpu(), tol=2e-6, what="multiclass step vs drop-in")
    assert eco is not None
