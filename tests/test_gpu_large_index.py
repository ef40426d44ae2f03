"""Tensors with more than 2**31 elements: every element offset past the first 8 GiB needs 64-bit index arithmetic in the
kernels.  The oracle cannot run at that size, so the tests use a size-independent property of the path (every
loss is a function of global sums over (N,H,W)): a tensor made of K copies of a small tile, stacked along H inside every
plane, has K times the tile's sums -- so its thresholded pixel counts are EXACTLY K times the tile's, its loss values equal
the tile's (the 1e-7 epsilons of the Dice ratios aside), and its gradient is the tile's gradient divided by K in every copy,
also in the copies that lie beyond element 2**31.  The tile itself is checked against the oracle (the reference's semantics
in eager torch ops on the same device)."""
import numpy as np
import pytest
import torch

from parity import assert_grad_close, assert_losses_close

pytestmark = pytest.mark.gpu

UP = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]
UP_ALL = [0.3, 1.0, 0.7, 0.2, 1.0, 0.5, 1.0]


def _need_gib(gib):
    free, _ = torch.cuda.mem_get_info()
    if free < gib * 2 ** 30:
        pytest.skip(f"needs {gib} GiB of free device memory")


def _combine(losses, up):
    return sum(float(w) * l for w, l in zip(up, losses) if w != 0.0)


def _check_copies(grad_big, grad_tile, k, tol, what):
    """grad_big [N,C,K*h,W] against grad_tile [N,C,h,W] / K in every copy, plane by plane (bounded temporaries)."""
    n, c, h, w = grad_tile.shape
    scale = float(grad_tile.abs().max()) / k
    worst = 0.0
    for i in range(n):
        for j in range(c):
            d = grad_big[i, j].view(k, h, w) - (grad_tile[i, j] / k)
            worst = max(worst, float(d.abs().max()) / scale)
            del d
    assert worst < tol, (what, worst)


def test_scoring_counts_beyond_2_31_elements():
    from ecologysemanticsegmentation_b200 import ops
    from oracle import counts as oc
    _need_gib(24)
    torch.manual_seed(61)
    n, c, h, w, k = 2, 2, 64, 1024, 8200                      # 2 * 2 * 8200 * 65536 = 2**31 + 2.1e6 elements
    z = (torch.randn(n, c, h, w) * 2).cuda()
    lab = (torch.rand(n, c, h, w) > 0.6).float().cuda()
    thr = torch.tensor([0.8, 0.9], dtype=torch.float32, device="cuda")
    tile_counts, tile_soft = ops.dice_counts(z, lab, thr)
    for t in range(2):
        assert (tile_counts[t].cpu().numpy() == oc.batch_counts(z, lab, float(thr[t]))).all()   # the tile against the oracle
    zb = z.repeat(1, 1, k, 1)
    lb = lab.to(torch.uint8).repeat(1, 1, k, 1)               # byte masks: 2 GiB instead of 8
    assert zb.numel() > 2 ** 31
    counts, soft = ops.dice_counts(zb, lb, thr)
    assert torch.equal(counts, tile_counts * k)
    np.testing.assert_allclose(soft.cpu().numpy(), tile_soft.cpu().numpy() * k, rtol=1e-7)   # (fp32 per-tile partials, folded in fp64)
    counts1, _ = ops.dice_counts(zb, lb, thr[:1])             # the one-threshold kernel
    assert torch.equal(counts1[0], tile_counts[0] * k)
    _, soft_u, _ = ops.dice_counts_ex(zb, lb, None, ununion_preds=True)   # (C == 2: the un-union leaves both channels alone)
    np.testing.assert_allclose(soft_u.cpu().numpy(), tile_soft.cpu().numpy() * k, rtol=1e-7)   # (fp32 per-tile partials, folded in fp64)


def test_fused_composite_step_beyond_2_31_elements():
    from ecologysemanticsegmentation_b200 import fused
    from oracle import torch_port as tp
    _need_gib(40)
    torch.manual_seed(62)
    n, h, w, k = 2, 64, 1024, 5500                            # 2 * 3 * 5500 * 65536 = 2.16e9 > 2**31 elements
    z = torch.randn(n, 3, h, w).cuda()
    u = torch.rand(n, 1, h, w)
    g = torch.cat([u < 0.5, u < 0.5 * 0.43197708, u < 0.5 * 0.22319692], 1).float().cuda()
    zr = z.clone().requires_grad_(True)
    np.random.seed(0)
    ref = tp.losses_composite(torch.sigmoid(zr), g, True)     # the tile against the oracle ...
    _combine(ref, UP).backward()
    np.random.seed(0)
    step = fused.CompositeLossStep(UP)
    l_tile, g_tile = step(z, g)
    assert_losses_close(l_tile.cpu().numpy(), ref, what="tile")
    assert_grad_close(g_tile.cpu(), zr.grad.cpu(), what="tile")
    zb = z.repeat(1, 1, k, 1)
    gb = g.to(torch.uint8).repeat(1, 1, k, 1)
    assert zb.numel() > 2 ** 31
    l_big, g_big = step(zb, gb)                               # ... and K copies of it against the tile
    assert_losses_close(l_big.cpu().numpy(), [float(v) for v in l_tile.cpu()], what="K copies")
    _check_copies(g_big, zr.grad, k, 1e-5, "K copies: gradient")


def test_leaf_step_beyond_2_31_elements_in_one_channel():
    """cfg1's kernels (C == 1) with more than 2**31 elements in the one channel: per-channel element counts and offsets
    beyond 32 bits."""
    from ecologysemanticsegmentation_b200 import fused
    from oracle import torch_port as tp
    _need_gib(40)
    torch.manual_seed(63)
    n, h, w, k = 3, 64, 1024, 11000                           # 3 * 11000 * 65536 = 2.16e9 > 2**31 elements
    z = torch.randn(n, 1, h, w).cuda()
    g = (torch.rand(n, 1, h, w) > 0.5).float().cuda()
    zr = z.clone().requires_grad_(True)
    ref = tp.losses_composite(torch.sigmoid(zr), g, False, 0.5)
    _combine(ref, UP_ALL).backward()
    step = fused.LeafLossStep(UP_ALL, doubling=2.0, background_weight=0.5)
    zb = z.repeat(1, 1, k, 1)
    gb = g.repeat(1, 1, k, 1)
    assert zb.numel() > 2 ** 31
    l_big, g_big = step(zb, gb)
    assert_losses_close(l_big.cpu().numpy(), ref, what="K copies")
    _check_copies(g_big, zr.grad, k, 1e-5, "K copies: gradient")
