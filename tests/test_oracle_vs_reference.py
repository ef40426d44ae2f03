"""Pins the oracle to the UNMODIFIED reference (only where /root/reference exists, i.e. the build container):
oracle.torch_port must be bit-identical to it on CPU, values and autograd gradients."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")

UP = [0.3, 1.0, 0.7, 0.2, 1.0, 0.5, 1.0]


def _run(fn, z, g, *args):
    zz = z.clone().requires_grad_(True)
    losses = fn(torch.sigmoid(zz), g, *args)
    sum(w * l for w, l in zip(UP, losses)).backward()
    return [float(v.detach()) for v in losses], zz.grad.clone()


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


@pytest.mark.parametrize("C,comp,bw,es", [(1, False, 0, False), (1, False, 0.5, False), (3, False, 0.5, False),
                                           (3, True, 0, False), (3, True, 0, True), (2, False, 0, False)])
def test_losses_fn_bit_identical(ref, C, comp, bw, es):
    from oracle import torch_port as tp
    lf, lc, tm = ref
    torch.manual_seed(0)
    z = torch.randn(3, C, 24, 20)
    g = (torch.rand(3, C, 24, 20) > 0.5).float()
    np.random.seed(7)
    a = _run(lc.losses_fn, z, g, comp, bw, es)
    np.random.seed(7)
    b = _run(tp.losses_composite, z, g, comp, bw, es)
    assert a[0] == b[0]
    assert torch.equal(a[1], b[1])


@pytest.mark.parametrize("C", [1, 3])
def test_train_multiclass_flavour_bit_identical(ref, C):
    from oracle import torch_port as tp
    lf, lc, tm = ref
    torch.manual_seed(1)
    z = torch.randn(2, C, 16, 16)
    g = (torch.rand(2, C, 16, 16) > 0.5).float()
    a = _run(tm, z, g, False, 0.5)
    b = _run(tp.losses_train_multiclass, z, g, False, 0.5)
    assert a[0] == b[0] and torch.equal(a[1], b[1])
    # lc == 2 x tm
    c = _run(lc.losses_fn, z, g, False, 0.5 if C == 1 else 0)
    np.testing.assert_allclose(np.array(c[0]), 2 * np.array(a[0]), rtol=1e-6)


def test_primitives_bit_identical(ref):
    from oracle import torch_port as tp
    lf, lc, tm = ref
    torch.manual_seed(2)
    a = torch.rand(2, 3, 8, 8)
    b = torch.rand(2, 3, 8, 8)
    pairs = [
        (lf.cross_entropy_loss(a, b, bce=True), tp.pair_bce(a, b)),
        (lf.cross_entropy_loss(a, b, background_weight=0.3), tp.pair_soft_ce(a, b, 0.3)),
        (lf.focal_loss(a, b, background_weight=0.4), tp.pair_focal(a, b, background_weight=0.4)),
        (lf.dice_loss(a, b), tp.pair_dice(a, b)),
        (lf.dice_loss(a, b, generalized=True), tp.pair_dice(a, b, generalized=True)),
        (lf.twersky_loss(a, b, background_weight=0.2), tp.pair_tversky(a, b, background_weight=0.2)),
        (lf.focal_dice_coefficient(a, b, background_weight=0.2), tp.pair_focal_dice(a, b, background_weight=0.2)),
    ]
    for r, o in pairs:
        assert float(r) == float(o)
    for r, o in zip(lf.classification_dice_loss(a, b), tp.pair_dice_family(a, b)):
        assert float(r) == float(o)


def test_eval_matches_reference_lines(ref):
    from oracle import torch_port as tp
    lf, lc, tm = ref
    torch.manual_seed(3)
    z = torch.randn(2, 3, 16, 16) * 2
    lab = (torch.rand(2, 3, 16, 16) > 0.6).float()
    for thr in (None, 0.8):
        out = torch.sigmoid(z)
        if thr is not None:
            out[out > thr] = 1
            out[out != 1] = 0
        want = [float(-lf.dice_loss(out[:, c:c + 1], lab[:, c:c + 1], background_weight=0)) for c in range(3)]
        got = [float(v) for v in tp.eval_batch_dice(z, lab, thr)]
        assert want == got


def test_adjacent_steps_bit_identical(ref):
    from oracle import torch_port as tp
    union_cls, union_bat, seq = ref_loader.load_adjacent()
    torch.manual_seed(4)
    ann = (torch.rand(4, 5, 6, 6) > 0.6).float()
    prob = torch.rand(4, 5, 6, 6)
    for ex in ([0], [0, 3], []):
        assert torch.equal(union_cls(ann.clone(), ex), tp.union_sets_descending(ann.clone(), ex))
        assert torch.equal(union_cls(prob.clone(), ex, reverse=True), tp.union_sets_descending(prob.clone(), ex, True))
        assert torch.equal(union_bat(ann.clone(), ex), tp.union_sets_descending_batchdim(ann.clone(), ex))
    z = torch.randn(3, 3, 8, 8)
    u = torch.rand(3, 1, 8, 8)
    g = torch.cat([u < 0.6, u < 0.4, u < 0.2], 1).float()   # nested, so g_1 - g_2 stays in {0, 1}
    a, b = _run(seq, z, g), _run(tp.losses_sequential_densenet, z, g)
    assert a[0] == b[0] and torch.equal(a[1], b[1])
    with pytest.raises(ValueError):
        seq(torch.rand(2, 1, 4, 4), torch.rand(2, 1, 4, 4), True)
    with pytest.raises(ValueError):
        tp.losses_sequential_densenet(torch.rand(2, 1, 4, 4), torch.rand(2, 1, 4, 4), True)
