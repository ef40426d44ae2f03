"""The oracle against the committed golden vectors (generated from the real reference by
tests/golden/make_golden.py).  Runs anywhere, no GPU."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_small.npz"))
META = json.load(open(os.path.join(HERE, "golden", "golden_meta.json")))
UP_ALL = [0.3, 1.0, 0.7, 0.2, 1.0, 0.5, 1.0]
UP_CFG2 = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]
RT = 2e-6  # CPU-vs-CPU, possibly different SIMD dispatch than the machine that wrote the fixtures


def _t(name):
    return torch.from_numpy(G[name].copy())


def _run(fn, x, g, up, *args, **kw):
    x = x.clone().requires_grad_(True)
    losses = fn(x, g, *args, **kw)
    sum(float(w) * l for w, l in zip(up, losses) if w != 0.0).backward()
    return np.array([float(v.detach()) for v in losses]), x.grad.numpy()


def test_torch_port_small_cases():
    from oracle import torch_port as tp
    p, gi, gn = _t("p"), _t("g_iid"), _t("g_nested")
    for bw in (0, 0.5):
        l, gr = _run(tp.losses_composite, p[:, :1], gi[:, :1], UP_ALL, False, bw)
        np.testing.assert_allclose(l, G[f"leaf_lc_bw{bw}_losses"], rtol=RT)
        np.testing.assert_allclose(gr, G[f"leaf_lc_bw{bw}_grad"], rtol=1e-4, atol=1e-9)
        l, gr = _run(tp.losses_train_multiclass, p[:, :1], gi[:, :1], UP_ALL, False, bw)
        np.testing.assert_allclose(l, G[f"leaf_tm_bw{bw}_losses"], rtol=RT)
    l, gr = _run(tp.losses_composite, p, gi, UP_ALL, False, 0.7)
    np.testing.assert_allclose(l, G["plain_lc_losses"], rtol=RT)
    np.testing.assert_allclose(gr, G["plain_lc_grad"], rtol=1e-4, atol=1e-9)
    for name, g in (("nested", gn), ("iid", gi)):
        np.random.seed(0)
        l, gr = _run(tp.losses_composite, p, g, UP_ALL, True)
        np.testing.assert_allclose(l, G[f"comp_{name}_losses"], rtol=RT)
        np.testing.assert_allclose(gr, G[f"comp_{name}_grad"], rtol=1e-4, atol=1e-8)
    np.random.seed(123)
    l, gr = _run(tp.losses_composite, p, gn, UP_CFG2, True, 0, True)
    np.testing.assert_allclose(l, G["comp_es_losses"], rtol=RT)
    assert (np.random.get_state()[1] == G["comp_es_rng_after"]).all()


def test_closed_form_small_cases():
    """The sufficient-statistics formulation (what the kernels implement) reproduces the reference's numbers."""
    from oracle import closed_form as cf, torch_port as tp
    p, gi, gn = G["p"].astype(np.float64), G["g_iid"].astype(np.float64), G["g_nested"].astype(np.float64)
    up = np.array(UP_ALL)
    for bw in (0, 0.5):
        s = cf.leaf_sums(p[:, :1], gi[:, :1])
        np.testing.assert_allclose(cf.leaf_losses(s, bw, 2.0), G[f"leaf_lc_bw{bw}_losses"], rtol=1e-5, atol=1e-12)
        ga, _ = cf.leaf_grad(p[:, :1], gi[:, :1], cf.leaf_coefs(s, 2.0 * up, bw))
        ref = G[f"leaf_lc_bw{bw}_grad"]
        assert np.abs(ga - ref).max() / np.abs(ref).max() < 1e-5
    l, gx = cf.plain_losses_and_grad(p, gi, up)
    np.testing.assert_allclose(l, G["plain_lc_losses"], rtol=1e-5, atol=1e-12)
    assert np.abs(gx - G["plain_lc_grad"]).max() / np.abs(G["plain_lc_grad"]).max() < 1e-5
    for name, g in (("nested", gn), ("iid", gi)):
        np.random.seed(0)
        w = cf.pair_weights(tp.DEFAULT_RATIOS)
        l, gx = cf.composite_losses_and_grad(p, g, w, up)
        np.testing.assert_allclose(l, G[f"comp_{name}_losses"], rtol=1e-5, atol=1e-12)
        ref = G[f"comp_{name}_grad"]
        assert np.abs(gx - ref).max() / np.abs(ref).max() < 1e-5


def test_primitives_against_golden():
    from oracle import torch_port as tp
    fns = {
        "bce": lambda a, b: tp.pair_bce(a, b),
        "softce": lambda a, b: tp.pair_soft_ce(a, b),
        "softce_bw": lambda a, b: tp.pair_soft_ce(a, b, 0.3),
        "focal": lambda a, b: tp.pair_focal(a, b),
        "focal_bw": lambda a, b: tp.pair_focal(a, b, factor=1, background_weight=0.4),
        "dice": lambda a, b: tp.pair_dice(a, b),
        "dice_bw0": lambda a, b: tp.pair_dice(a, b, background_weight=0),
        "gdice": lambda a, b: tp.pair_dice(a, b, generalized=True, background_weight=0.5),
        "twersky": lambda a, b: tp.pair_tversky(a, b, background_weight=0.25),
        "focal_dice": lambda a, b: tp.pair_focal_dice(a, b, background_weight=0.25),
        "cls_dice": lambda a, b: tp.pair_dice_family(a, b),
        "cls_dice_f10": lambda a, b: tp.pair_dice_family(a, b, factor=10, background_weight=0),
    }
    for name, fn in fns.items():
        a = _t("prim_a").requires_grad_(True)
        b = _t("prim_b").requires_grad_(True)
        out = fn(a, b)
        outs = out if isinstance(out, tuple) else (out,)
        sum((k + 1.0) * o for k, o in enumerate(outs)).backward()
        np.testing.assert_allclose(np.array([float(o.detach()) for o in outs]), G[f"prim_{name}_val"], rtol=RT, err_msg=name)
        if a.grad is not None:
            np.testing.assert_allclose(a.grad.numpy(), G[f"prim_{name}_ga"], rtol=1e-4, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(b.grad.numpy(), G[f"prim_{name}_gb"], rtol=1e-4, atol=1e-9, err_msg=name)


def test_eval_against_golden():
    from oracle import counts as oc, torch_port as tp
    z, lab = _t("eval_z"), _t("eval_lab")
    for thr in (None, 0.8, 0.9):
        d = np.array([float(v) for v in tp.eval_batch_dice(z, lab, thr)])
        np.testing.assert_allclose(d, G[f"eval_dice_{thr}"], rtol=RT)
        if thr is not None:
            c = oc.batch_counts(z, lab, thr)
            # CPU sigmoid implementations may differ by one ulp right at the threshold on another machine
            assert np.abs(c - G[f"eval_counts_{thr}"]).max() <= 1
            np.testing.assert_allclose(oc.dice_from_counts(G[f"eval_counts_{thr}"]), G[f"eval_dice_{thr}"], rtol=1e-6)


def test_survey_known_answers():
    """SURVEY.md 8(c) KATs, recomputed by the oracle from the same seed."""
    from oracle import torch_port as tp
    kat = META["cases"]["kat_seed0_4x3x64x64"]
    torch.manual_seed(0)
    p = torch.sigmoid(torch.randn(4, 3, 64, 64))
    g = (torch.rand(4, 3, 64, 64) > 0.5).float()
    if hashlib.sha256(g.numpy().tobytes()).hexdigest() != kat["g_sha"]:
        pytest.skip("torch RNG stream differs from the fixture machine")
    np.testing.assert_allclose([float(v) for v in tp.losses_composite(p[:, :1], g[:, :1])], kat["lc_c1"], rtol=RT)
    np.testing.assert_allclose([float(v) for v in tp.losses_composite(p, g)], kat["lc_c3"], rtol=RT)
    np.random.seed(0)
    np.testing.assert_allclose([float(v) for v in tp.losses_composite(p, g, True)], kat["lc_c3_composite"], rtol=RT)
    np.testing.assert_allclose([float(v) for v in tp.losses_train_multiclass(p, g)], kat["tm_c3"], rtol=RT)
    # the survey's printed values
    np.testing.assert_allclose(kat["lc_c1"], [0, 1.51038, 16.11810, -3.28321, -1.64160, -3.65241, 1.33557], rtol=1e-5)
    np.testing.assert_allclose(kat["lc_c3_composite"], [0, 111.96028, 462.16394, -222.08194, -111.04097, -215.71408, 201.62666], rtol=1e-5)


def test_adjacent_steps_against_golden():
    """Label union / un-union (utils/subsets_union.py:8-32, train_multiclass.py:32-45) and the sequential-variant
    losses_fn (train_multiclass_sequential_densenetloss.py:272-362)."""
    from oracle import torch_port as tp
    ann, prob = _t("union_ann"), _t("union_prob")
    for tag, ex in (("e0", [0]), ("e02", [0, 2]), ("none", [])):
        assert np.array_equal(tp.union_sets_descending(ann.clone(), ex).numpy(), G[f"union_cls_fwd_{tag}"])
        np.testing.assert_allclose(tp.union_sets_descending(prob.clone(), ex, True).numpy(), G[f"union_cls_rev_{tag}"], rtol=1e-6)
        assert np.array_equal(tp.union_sets_descending_batchdim(ann.clone(), ex).numpy(), G[f"union_bat_fwd_{tag}"])
    l, gr = _run(tp.losses_sequential_densenet, _t("p"), _t("g_nested"), UP_ALL)
    np.testing.assert_allclose(l, G["seq_losses"], rtol=RT)
    np.testing.assert_allclose(gr, G["seq_grad"], rtol=1e-4, atol=1e-9)
    l, gr = _run(tp.losses_sequential_densenet, _t("p")[:, :1], _t("g_iid")[:, :1], UP_ALL, False, 0.5)
    np.testing.assert_allclose(l, G["seq_c1_losses"], rtol=RT)


def test_torch_port_primitives_with_nondefault_keywords():
    """The oracle's parametric primitives against the reference's outputs for non-default gamma / alpha / beta
    (tests/golden/make_golden_shaped.py)."""
    from oracle import torch_port as tp
    GS = np.load(os.path.join(HERE, "golden", "golden_shaped.npz"))
    cases = {
        "focal_g2": lambda a, b: tp.pair_focal(a, b, gamma=2.0),
        "focal_g07_bw": lambda a, b: tp.pair_focal(a, b, gamma=0.7, factor=1, background_weight=0.4),
        "focal_g3_bw": lambda a, b: tp.pair_focal(a, b, gamma=3, factor=0.5, background_weight=1),
        "twersky_a7b3": lambda a, b: tp.pair_tversky(a, b, alpha=0.7, beta=0.3),
        "twersky_a2b8_bw": lambda a, b: tp.pair_tversky(a, b, alpha=0.2, beta=0.8, background_weight=0.25),
        "focal_dice_g1": lambda a, b: tp.pair_focal_dice(a, b, gamma=1.0),
        "focal_dice_g25_bw": lambda a, b: tp.pair_focal_dice(a, b, gamma=2.5, background_weight=0.25),
    }
    for name, fn in cases.items():
        a, b = _t("prim_a").requires_grad_(True), _t("prim_b").requires_grad_(True)
        v = fn(a, b)
        v.backward()
        np.testing.assert_allclose(float(v.detach()), GS[f"{name}_val"][0], rtol=RT, err_msg=name)
        np.testing.assert_allclose(b.grad.numpy(), GS[f"{name}_gb"], rtol=1e-4, atol=1e-9, err_msg=name)
        if a.grad is not None:
            np.testing.assert_allclose(a.grad.numpy(), GS[f"{name}_ga"], rtol=1e-4, atol=1e-9, err_msg=name)


def test_masks_u8_known_answers():
    """test_multiclass.py:90-92 `(t.numpy() * 255).astype(np.uint8)`: float32 product, truncation toward zero."""
    from oracle import counts as oc
    from oracle import torch_port as tp
    t = torch.tensor([0.0, 1.0, 0.5, 0.999, 0.003921569, 0.0039, 0.99999994], dtype=torch.float32)
    assert oc.masks_u8(t).tolist() == [0, 255, 127, 254, 1, 0, 254]
    assert oc.masks_u8(t).dtype == np.uint8
    z = _t("eval_z")
    out = tp.threshold_inplace(torch.sigmoid(z), 0.8)
    m = oc.masks_u8(out)
    assert set(np.unique(m).tolist()) <= {0, 255}
    assert int((m == 255).sum()) == int(G["eval_counts_0.8"][:, 1].sum())   # |out| counts of the golden file


def test_sequential_test_scoring_against_golden():
    """ess/test_multiclass_sequential_densenetloss.py:62,66,97-99 (sigmoid -> prediction un-union -> per-class soft Dice): the
    oracle's restatement against outputs of the reference's own functions (tests/golden/make_golden_seqtest.py)."""
    import os
    from oracle import torch_port as tp
    S = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_seqtest.npz"))
    for tag in ("c3", "c4", "c2"):
        z, lab = torch.from_numpy(S[f"{tag}_z"]), torch.from_numpy(S[f"{tag}_lab"])
        out = tp.union_sets_descending(torch.sigmoid(z), reverse=True)
        assert np.array_equal(out.numpy(), S[f"{tag}_out"])
        dice = [-float(tp.pair_dice(out[:, c:c + 1], lab[:, c:c + 1], background_weight=0)) for c in range(lab.shape[1])]
        np.testing.assert_allclose(dice, S[f"{tag}_dice"], rtol=1e-7)
