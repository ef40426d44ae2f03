"""CUDA path against the committed golden vectors (outputs of the UNMODIFIED reference, see
tests/golden/make_golden.py) and against size-independent properties at BASELINE.json's full sizes."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from parity import assert_grad_close, assert_losses_close, grad_errors, TOL, TOL_BF16

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_small.npz"))
META = json.load(open(os.path.join(HERE, "golden", "golden_meta.json")))
UP_ALL = [0.3, 1.0, 0.7, 0.2, 1.0, 0.5, 1.0]
UP_CFG2 = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]
UP_CFG1 = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 0.0]


def _c(name):
    return torch.from_numpy(G[name].copy()).cuda()


def _run(fn, x, g, up, *args, **kw):
    x = x.clone().requires_grad_(True)
    losses = fn(x, g, *args, **kw)
    assert len(losses) == 7
    sum(float(w) * l for w, l in zip(up, losses) if w != 0.0).backward()
    return [float(v) for v in losses], x.grad


def _sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def test_small_cases_vs_reference_outputs():
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import train_multiclass as tm
    p, gi, gn = _c("p"), _c("g_iid"), _c("g_nested")
    for bw in (0, 0.5):
        l, gr = _run(eco.losses_fn, p[:, :1], gi[:, :1], UP_ALL, False, bw)
        assert_losses_close(l, G[f"leaf_lc_bw{bw}_losses"], what=f"leaf lc bw={bw}")
        assert_grad_close(gr.cpu(), G[f"leaf_lc_bw{bw}_grad"], what=f"leaf lc bw={bw}")
        l, gr = _run(tm.losses_fn, p[:, :1], gi[:, :1], UP_ALL, False, bw)
        assert isinstance(l, list)
        assert_losses_close(l, G[f"leaf_tm_bw{bw}_losses"], what=f"leaf tm bw={bw}")
        assert_grad_close(gr.cpu(), G[f"leaf_tm_bw{bw}_grad"], what=f"leaf tm bw={bw}")
    l, gr = _run(eco.losses_fn, p, gi, UP_ALL, False, 0.7)
    assert_losses_close(l, G["plain_lc_losses"], what="plain lc")
    assert_grad_close(gr.cpu(), G["plain_lc_grad"], what="plain lc")
    l, gr = _run(tm.losses_fn, p, gi, UP_ALL, True, 0.7)
    assert_losses_close(l, G["plain_tm_losses"], what="plain tm")
    assert_grad_close(gr.cpu(), G["plain_tm_grad"], what="plain tm")
    for name, g in (("nested", gn), ("iid", gi)):
        np.random.seed(0)
        l, gr = _run(eco.losses_fn, p, g, UP_ALL, True)
        assert_losses_close(l, G[f"comp_{name}_losses"], what=f"composite {name}")
        assert_grad_close(gr.cpu(), G[f"comp_{name}_grad"], what=f"composite {name}")
    np.random.seed(123)
    l, gr = _run(eco.losses_fn, p, gn, UP_CFG2, True, 0, True)
    assert (np.random.get_state()[1] == G["comp_es_rng_after"]).all()
    assert_losses_close(l, G["comp_es_losses"], what="composite early_stopped")
    assert_grad_close(gr.cpu(), G["comp_es_grad"], what="composite early_stopped")


def test_train_multiclass_c1_composite_raises_like_reference():
    from ecologysemanticsegmentation_b200 import train_multiclass as tm
    with pytest.raises(ValueError):
        tm.losses_fn(_c("p")[:, :1], _c("g_iid")[:, :1], True)


PRIMS = {
    "bce": lambda lf, a, b: lf.cross_entropy_loss(a, b, bce=True),
    "softce": lambda lf, a, b: lf.cross_entropy_loss(a, b),
    "softce_bw": lambda lf, a, b: lf.cross_entropy_loss(a, b, background_weight=0.3),
    "focal": lambda lf, a, b: lf.focal_loss(a, b),
    "focal_bw": lambda lf, a, b: lf.focal_loss(a, b, factor=1, background_weight=0.4),
    "dice": lambda lf, a, b: lf.dice_loss(a, b),
    "dice_bw0": lambda lf, a, b: lf.dice_loss(a, b, background_weight=0),
    "gdice": lambda lf, a, b: lf.dice_loss(a, b, generalized=True, background_weight=0.5),
    "twersky": lambda lf, a, b: lf.twersky_loss(a, b, background_weight=0.25),
    "focal_dice": lambda lf, a, b: lf.focal_dice_coefficient(a, b, background_weight=0.25),
    "cls_dice": lambda lf, a, b: lf.classification_dice_loss(a, b),
    "cls_dice_f10": lambda lf, a, b: lf.classification_dice_loss(a, b, factor=10, background_weight=0),
}


@pytest.mark.parametrize("name", sorted(PRIMS))
def test_primitives_vs_reference_outputs(name):
    """Every function of loss_functions.py on two continuous tensors, gradients w.r.t. BOTH slots."""
    from ecologysemanticsegmentation_b200 import loss_functions as lf
    a = _c("prim_a").requires_grad_(True)
    b = _c("prim_b").requires_grad_(True)
    out = PRIMS[name](lf, a, b)
    outs = out if isinstance(out, tuple) else (out,)
    sum((k + 1.0) * o for k, o in enumerate(outs)).backward()
    assert_losses_close([float(o) for o in outs], G[f"prim_{name}_val"], what=name)
    ref_a, ref_b = G[f"prim_{name}_ga"], G[f"prim_{name}_gb"]
    if np.abs(ref_a).max() > 0:
        assert_grad_close(a.grad.cpu(), ref_a, what=name + " d/da")
    else:
        assert a.grad is None or float(a.grad.abs().max()) == 0.0
    assert_grad_close(b.grad.cpu(), ref_b, what=name + " d/db")


GS = np.load(os.path.join(HERE, "golden", "golden_shaped.npz"))
SHAPED = {
    "focal_g2": lambda lf, a, b: lf.focal_loss(a, b, gamma=2.0),
    "focal_g07_bw": lambda lf, a, b: lf.focal_loss(a, b, gamma=0.7, factor=1, background_weight=0.4),
    "focal_g3_bw": lambda lf, a, b: lf.focal_loss(a, b, gamma=3, factor=0.5, background_weight=1),
    "twersky_a7b3": lambda lf, a, b: lf.twersky_loss(a, b, alpha=0.7, beta=0.3),
    "twersky_a2b8_bw": lambda lf, a, b: lf.twersky_loss(a, b, alpha=0.2, beta=0.8, background_weight=0.25),
    "focal_dice_g1": lambda lf, a, b: lf.focal_dice_coefficient(a, b, gamma=1.0),
    "focal_dice_g25_bw": lambda lf, a, b: lf.focal_dice_coefficient(a, b, gamma=2.5, background_weight=0.25),
}


@pytest.mark.parametrize("name", sorted(SHAPED))
def test_primitives_with_nondefault_keywords_vs_reference_outputs(name):
    """focal_loss(gamma=), twersky_loss(alpha=, beta=), focal_dice_coefficient(gamma=): golden_shaped.npz holds
    the unmodified reference's values and gradients (tests/golden/make_golden_shaped.py)."""
    from ecologysemanticsegmentation_b200 import loss_functions as lf
    a = _c("prim_a").requires_grad_(True)
    b = _c("prim_b").requires_grad_(True)
    out = SHAPED[name](lf, a, b)
    out.backward()
    assert_losses_close([float(out)], GS[f"{name}_val"], what=name)
    ref_a, ref_b = GS[f"{name}_ga"], GS[f"{name}_gb"]
    if np.abs(ref_a).max() > 0:
        assert_grad_close(a.grad.cpu(), ref_a, what=name + " d/da")
    else:
        assert a.grad is None or float(a.grad.abs().max()) == 0.0
    assert_grad_close(b.grad.cpu(), ref_b, what=name + " d/db")


def test_nondefault_focal_gamma_edge_values_match_torch_pow():
    """b exactly 0 and 1 (labels in the pred slot) and an integer exponent: powf semantics = torch.pow's."""
    from ecologysemanticsegmentation_b200 import loss_functions as lf
    from oracle import torch_port as tp
    torch.manual_seed(9)
    b = (torch.rand(2, 1, 16, 16) > 0.5).float()
    b[0, 0, 0, :8] = torch.rand(8)
    a = torch.rand(2, 1, 16, 16)
    for gamma, bw in ((2.0, 0.0), (2.0, 0.5), (1.0, 0.3)):
        ref = tp.pair_focal(a, b, gamma=gamma, factor=1, background_weight=bw)
        ours = lf.focal_loss(a.cuda(), b.cuda(), gamma=gamma, factor=1, background_weight=bw)
        assert_losses_close([float(ours)], [float(ref)], what=f"focal gamma={gamma} bw={bw}")


def test_binary_cross_entropy_list_sums_through_cpu_buffer():
    from ecologysemanticsegmentation_b200 import loss_functions as lf
    a, b = _c("prim_a"), _c("prim_b")
    out = lf.binary_cross_entropy_list([a, a[:, :1]], [b, b[:, :1]])
    assert out.device.type == "cpu"
    ref = float(lf.cross_entropy_loss(a, b, bce=True)) + float(lf.cross_entropy_loss(a[:, :1], b[:, :1], bce=True))
    assert abs(float(out) - ref) / abs(ref) < 1e-6
    with pytest.raises(TypeError):
        lf.cross_entropy_list([a], [b])


def test_eval_small_vs_reference_outputs():
    from ecologysemanticsegmentation_b200 import test_multiclass as tmc
    z, lab = _c("eval_z"), _c("eval_lab")
    d = tmc.score_batch(z, lab)
    assert_losses_close(d.cpu().numpy(), G["eval_dice_None"], what="soft dice")
    for thr in (0.8, 0.9):
        d, counts, _ = tmc.score_batch(z, lab, thr, return_counts=True)
        assert (counts[0].cpu().numpy() == G[f"eval_counts_{thr}"]).all()
        assert_losses_close(d.cpu().numpy(), G[f"eval_dice_{thr}"], what=f"dice@{thr}")


@pytest.mark.parametrize("bw", [0, 0.5])
def test_cfg1_full_size_vs_reference_digest(bw):
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import train_multiclass as tm
    from ecologysemanticsegmentation_b200.synthetic import make_config
    z, g = make_config("cfg1")
    ref = META["cases"][f"cfg1_lc_bw{bw}"]
    p = torch.sigmoid(z)
    if _sha(p) != ref["p_sha"] or _sha(g) != ref["g_sha"]:
        pytest.skip("synthetic generator produced different bits than on the fixture machine")
    for fn, case in ((eco.losses_fn, f"cfg1_lc_bw{bw}"), (tm.losses_fn, f"cfg1_tm_bw{bw}")):
        c = META["cases"][case]
        l, gr = _run(fn, p.cuda(), g.cuda(), UP_CFG1, False, bw)
        assert_losses_close(l, c["losses"], what=case)
        _check_digest(gr, c["grad_wrt_p"], case)


def _check_digest(grad, dig, what):
    flat = grad.double().flatten().cpu()
    idx = torch.tensor(dig["sample_idx"])
    err = (flat[idx] - torch.tensor(dig["sample_val"], dtype=torch.float64)).abs().max() / dig["max_abs"]
    assert float(err) <= TOL, f"{what}: sampled gradient max-norm error {float(err):.3e}"
    assert abs(float(flat.norm()) - dig["l2"]) / dig["l2"] <= TOL, what
    assert abs(float(flat.abs().max()) - dig["max_abs"]) / dig["max_abs"] <= TOL, what


def test_cfg2_full_size_vs_reference_digest():
    """BASELINE configs[1]: the headline workload, against the reference's own CPU fp32 outputs."""
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200.synthetic import make_config
    z, g = make_config("cfg2")
    c = META["cases"]["cfg2_composite"]
    p = torch.sigmoid(z)
    if _sha(p) != c["p_sha"] or _sha(g) != c["g_sha"]:
        pytest.skip("synthetic generator produced different bits than on the fixture machine")
    np.random.seed(c["np_seed"])
    l, gr = _run(eco.losses_fn, p.cuda(), g.cuda(), UP_CFG2, True)
    assert_losses_close(l, c["losses"], what="cfg2 composite")
    _check_digest(gr, c["grad_wrt_p"], "cfg2 composite")
    l, gr = _run(eco.losses_fn, p.cuda(), g.cuda(), UP_CFG2, False)
    assert_losses_close(l, META["cases"]["cfg2_plain"]["losses"], what="cfg2 plain")
    _check_digest(gr, META["cases"]["cfg2_plain"]["grad_wrt_p"], "cfg2 plain")
    # from logits, fused single launch: losses must agree as well (sigmoid differs from the CPU's by <= 1 ulp)
    from ecologysemanticsegmentation_b200.fused import CompositeLossStep
    np.random.seed(c["np_seed"])
    losses, dz = CompositeLossStep(UP_CFG2)(z.cuda(), g.cuda())
    assert_losses_close(losses.cpu().numpy(), c["losses"], what="cfg2 fused from logits")
    dig = c["grad_wrt_z_cpu_sigmoid"]
    flat = dz.double().flatten().cpu()
    assert abs(float(flat.norm()) - dig["l2"]) / dig["l2"] <= TOL


def test_cfg3_scoring_counts_and_properties():
    """Scoring at IMG 1024: golden (CPU-sigmoid) counts within a few ulp-ties, exact vs the same-device oracle,
    and additivity of counts over batch shards."""
    from ecologysemanticsegmentation_b200 import ops, test_multiclass as tmc
    from ecologysemanticsegmentation_b200.synthetic import make_inputs
    from oracle import counts as oc
    z, g = make_inputs(8, 3, 1024, 103)
    c = META["cases"]["cfg3_n8"]
    same_bits = _sha(z) == c["z_sha"] and _sha(g) == c["g_sha"]
    zc, gc = z.cuda(), g.cuda()
    for thr in (0.8, 0.9):
        d, counts, _ = tmc.score_batch(zc, gc, thr, return_counts=True)
        got = counts[0].cpu().numpy()
        assert (got == oc.batch_counts(zc, gc, thr)).all(), "not bit-exact vs the reference ops on this device"
        if same_bits:
            assert np.abs(got - np.array(c[f"counts_{thr}"])).max() <= 3  # CPU vs CUDA sigmoid ties
            assert_losses_close(d.cpu().numpy(), c[f"dice_{thr}"], what=f"cfg3 dice@{thr}")
        halves = [ops.dice_counts(zc[:4], gc[:4], torch.tensor([thr], device="cuda"))[0],
                  ops.dice_counts(zc[4:], gc[4:], torch.tensor([thr], device="cuda"))[0]]
        assert torch.equal(halves[0] + halves[1], counts)
    if same_bits:
        assert_losses_close(tmc.score_batch(zc, gc).cpu().numpy(), c["dice_None"], what="cfg3 soft dice")


def test_label_union_and_unun_vs_reference_outputs():
    """utils/subsets_union.py:8-32 (class dim, forward + reverse) and train_multiclass.py:32-45 (batch-dim twin)."""
    from ecologysemanticsegmentation_b200 import subsets_union as su, train_multiclass as tm
    for tag, ex in (("e0", [0]), ("e02", [0, 2]), ("none", [])):
        a = _c("union_ann")
        out = su.return_union_sets_descending_order(a, ex)
        assert out.data_ptr() == a.data_ptr(), "must work in place like the reference"
        assert np.array_equal(a.cpu().numpy(), G[f"union_cls_fwd_{tag}"])
        pr = _c("union_prob")
        su.return_union_sets_descending_order(pr, ex, reverse=True)
        np.testing.assert_allclose(pr.cpu().numpy(), G[f"union_cls_rev_{tag}"], rtol=1e-6)
        b = _c("union_ann")
        tm.return_union_sets_descending_order(b, ex)
        assert np.array_equal(b.cpu().numpy(), G[f"union_bat_fwd_{tag}"])
    # full-size property: forward union then reverse un-union gives back nested binary labels
    from ecologysemanticsegmentation_b200.synthetic import make_config
    _, g = make_config("cfg2")
    parts = torch.stack([g[:, 0], g[:, 1] - g[:, 2], g[:, 2]], 1).contiguous().cuda()   # disjoint ventral / dorsal
    u = su.return_union_sets_descending_order(parts.clone())
    assert torch.equal(u, g.cuda())
    back = su.return_union_sets_descending_order(u.clone(), reverse=True)
    assert torch.equal(back[:, 1:], parts[:, 1:])


def test_sequential_variant_losses_fn_vs_reference_outputs():
    from ecologysemanticsegmentation_b200 import train_multiclass_sequential_densenetloss as seq
    l, gr = _run(seq.losses_fn, _c("p"), _c("g_nested"), UP_ALL)
    assert_losses_close(l, G["seq_losses"], what="sequential C=3")
    assert_grad_close(gr.cpu(), G["seq_grad"], what="sequential C=3")
    l, gr = _run(seq.losses_fn, _c("p")[:, :1], _c("g_iid")[:, :1], UP_ALL, False, 0.5)
    assert_losses_close(l, G["seq_c1_losses"], what="sequential C=1")
    assert_grad_close(gr.cpu(), G["seq_c1_grad"], what="sequential C=1")
    with pytest.raises(ValueError):
        seq.losses_fn(_c("p")[:, :1], _c("g_iid")[:, :1], True)


def test_sequential_test_scoring_vs_reference_outputs():
    """ess/test_multiclass_sequential_densenetloss.py:62,66,97-99 -- the fused sigmoid -> un-union -> soft Dice read against
    outputs of the reference's own functions (tests/golden/make_golden_seqtest.py), 1e-5 relative (CPU vs CUDA sigmoid)."""
    from ecologysemanticsegmentation_b200 import test_multiclass_sequential_densenetloss as seq
    S = np.load(os.path.join(HERE, "golden", "golden_seqtest.npz"))
    for tag in ("c3", "c4", "c2"):
        z, lab = torch.from_numpy(S[f"{tag}_z"]).cuda(), torch.from_numpy(S[f"{tag}_lab"]).cuda()
        d = seq.score_batch(z, lab)
        np.testing.assert_allclose(d.cpu().numpy(), S[f"{tag}_dice"], rtol=1e-5)
        d8 = seq.score_batch(z, lab.to(torch.uint8))
        assert torch.equal(d8, d)
