"""Out-of-bounds writes: every kernel that writes a caller-provided buffer gets that buffer as the MIDDLE of a larger one whose
borders hold a canary; the borders must come back untouched and the result must equal the one written into a tight buffer.
Shapes end in partial tiles (H*W = 1296 = one 1024-pixel tile + 272) or are not 16-byte aligned at all (17 x 19)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

UP = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]
CANARY = -12345.5
PAD = 4096   # elements on either side


def _guarded(shape, dtype=torch.float32):
    numel = int(np.prod(shape))
    big = torch.full((numel + 2 * PAD,), CANARY, dtype=dtype, device="cuda")
    return big, big[PAD:PAD + numel].view(shape)


def _borders_intact(big):
    return bool((big[:PAD] == CANARY).all()) and bool((big[-PAD:] == CANARY).all())


@pytest.mark.parametrize("shape", [(5, 3, 36, 36), (2, 3, 16, 16), (3, 3, 17, 19), (1, 3, 4, 4)])
def test_step_gradients_stay_inside_their_buffers(shape):
    from ecologysemanticsegmentation_b200 import fused
    torch.manual_seed(71)
    z = torch.randn(shape).cuda()
    g = (torch.rand(shape) > 0.5).float().cuda()
    for make in (lambda: fused.CompositeLossStep(UP), lambda: fused.MulticlassLossStep(UP), lambda: fused.LeafLossStep(UP)):
        np.random.seed(0)
        try:
            l_ref, d_ref = make()(z, g)
        except Exception as exc:   # a step object that does not serve the shape must say so, not write anywhere
            assert "aligned" in str(exc) or "H*W" in str(exc) or "planes" in str(exc), exc
            continue
        big, out = _guarded(shape)
        np.random.seed(0)
        l, d = make()(z, g, out=out)
        assert d.data_ptr() == out.data_ptr()
        assert _borders_intact(big)
        assert torch.equal(d, d_ref) and torch.equal(l, l_ref)


@pytest.mark.parametrize("shape", [(5, 3, 36, 36), (3, 4, 17, 19)])
def test_inplace_union_and_inputs_in_the_middle_of_a_buffer(shape):
    """The in-place kernels on a view in the middle of a larger buffer, and the read-only kernels on inputs that start at an
    odd offset of their allocation (element offset 4096 + 1: 4-byte aligned only)."""
    from ecologysemanticsegmentation_b200 import ops, subsets_union
    torch.manual_seed(72)
    lab = (torch.rand(shape) > 0.6).float().cuda()
    ref = subsets_union.return_union_sets_descending_order(lab.clone())
    big, view = _guarded(shape)
    view.copy_(lab)
    subsets_union.return_union_sets_descending_order(view)
    assert _borders_intact(big) and torch.equal(view, ref)
    z = (torch.randn(shape) * 2).cuda()
    numel = z.numel()
    zb = torch.full((numel + 2 * PAD + 1,), CANARY, device="cuda")
    lb = torch.full((numel + 2 * PAD + 1,), CANARY, device="cuda")
    zo, lo = zb[PAD + 1:PAD + 1 + numel].view(shape), lb[PAD + 1:PAD + 1 + numel].view(shape)
    zo.copy_(z); lo.copy_(lab)
    thr = torch.tensor([0.8], dtype=torch.float32, device="cuda")
    for t in (None, thr, torch.tensor(np.arange(0.8, 0.99, 0.01), dtype=torch.float32, device="cuda")):
        c0, s0 = ops.dice_counts(z, lab, t)
        c1, s1 = ops.dice_counts(zo, lo, t)
        assert torch.equal(c0, c1)
        np.testing.assert_allclose(s1.cpu().numpy(), s0.cpu().numpy(), rtol=1e-6)
    m0 = ops.masks_u8(z, 0.8, False)
    m1 = ops.masks_u8(zo, 0.8, False)
    assert torch.equal(m0, m1)
