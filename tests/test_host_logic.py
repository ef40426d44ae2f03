"""Host-side mirror of the reference interface: names, signatures, LossList semantics, RNG-stream parity,
error behaviour, sharding helpers.  No GPU."""
import inspect

import numpy as np
import pytest
import torch


def test_public_names_match_reference_modules():
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import loss_composite, loss_functions, test_multiclass, train_multiclass
    for name in ("binary_cross_entropy_list", "cross_entropy_list", "focal_list", "classification_dice_list",
                 "cross_entropy_loss", "focal_loss", "dice_loss", "twersky_loss", "focal_dice_coefficient",
                 "classification_dice_loss"):
        assert hasattr(loss_functions, name), name
    for name in ("LossList", "losses_fn", "intersection_loss", "union_loss"):
        assert hasattr(loss_composite, name), name
    assert callable(eco.binary_cross_entropy)
    assert hasattr(train_multiclass, "losses_fn") and hasattr(test_multiclass, "test")


def _params(fn):
    return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()
            if p.kind != inspect.Parameter.KEYWORD_ONLY]


def test_signatures_match_reference():
    """Positional names and defaults as in loss_functions.py:26,46,52,82,96,110, loss_composite.py:21,87,92,
    train_multiclass.py:253, test_multiclass.py:30."""
    from ecologysemanticsegmentation_b200 import loss_composite as lc, loss_functions as lf, test_multiclass as tmc, \
        train_multiclass as tm
    E = inspect.Parameter.empty
    assert _params(lf.cross_entropy_loss) == [("gt", E), ("pred", E), ("weight", 0.3), ("bce", False), ("background_weight", 0)]
    assert _params(lf.focal_loss) == [("gt", E), ("pred", E), ("gamma", 1.5), ("factor", 0.1), ("background_weight", 0)]
    assert _params(lf.dice_loss) == [("gt", E), ("pred", E), ("generalized", False), ("background_weight", 1)]
    assert _params(lf.twersky_loss) == [("gt", E), ("pred", E), ("alpha", 0.5), ("beta", 0.3), ("background_weight", 0)]
    assert _params(lf.focal_dice_coefficient) == [("gt", E), ("pred", E), ("alpha", 0.5), ("beta", 0.3), ("gamma", 1.8), ("background_weight", 0)]
    assert _params(lf.classification_dice_loss) == [("gt", E), ("pred", E), ("factor", 1e3), ("background_weight", 1)]
    assert _params(lc.losses_fn) == [("x", E), ("g", E), ("composite_set_theory", False), ("background_weight", 0),
                                      ("early_stopped", False), ("relative_set_ratios", [1., 0.43197708, 0.22319692])]
    assert _params(lc.intersection_loss) == [("superset_p", E), ("set_p", E), ("set_g", E)]
    assert _params(lc.union_loss) == [("superset_p", E), ("set_p", E), ("superset_g", E)]
    assert _params(tm.losses_fn) == [("x", E), ("g", E), ("composite_set_theory", False), ("background_weight", 0), ("early_stopped", False)]
    assert _params(tmc.test)[:7] == [("net", E), ("dataloader", E), ("models_dir", "models/vgg"), ("results_dir", "test_results/"),
                                     ("batch_size", 1), ("saved_epoch", -1), ("single_model", False)]


def test_losslist_semantics():
    """loss_composite.py:9-17."""
    from ecologysemanticsegmentation_b200 import LossList
    a = LossList([1.0, 2.0])
    b = a
    a += LossList([10.0, 20.0])
    assert isinstance(a, LossList) and list(a) == [11.0, 22.0]
    assert list(b) == [1.0, 2.0], "+= must build a new list like the reference"
    with pytest.raises(AssertionError, match="same length"):
        a += LossList([1.0])
    assert list(a * 2.0) == [22.0, 44.0] and isinstance(a * 2.0, LossList)
    assert list(a * np.float64(0.5)) == [5.5, 11.0]
    for bad in (2, torch.tensor(2.0)):
        with pytest.raises(AssertionError, match="numerical weights"):
            a * bad


def test_pair_weight_draw_uses_the_reference_rng_stream():
    from ecologysemanticsegmentation_b200.loss_composite import DEFAULT_RATIOS, composite3_leaf_scales, draw_pair_weights
    from oracle import torch_port as tp
    for es in (False, True):
        np.random.seed(42)
        ours = draw_pair_weights(DEFAULT_RATIOS, es)
        st_ours = np.random.get_state()[1].copy()
        np.random.seed(42)
        ref = [(i, j) + tp.draw_pair_weights(DEFAULT_RATIOS, i, j, es) for i in range(2) for j in range(i + 1, 3)]
        st_ref = np.random.get_state()[1].copy()
        assert ours == ref and (st_ours == st_ref).all()
    np.random.seed(0)
    w = draw_pair_weights(DEFAULT_RATIOS, False)
    # pair(0,1) w_i=1, w_j=2.31494, w_d=1.76050
    np.testing.assert_allclose(w[0][2:], [1.0, 2.31494, 1.76050], rtol=1e-5)
    np.testing.assert_allclose(w[2][2:], [2.31494, 4.48035, 4.78973], rtol=1e-5)
    sc = composite3_leaf_scales(w)
    assert len(sc) == 21 and sc[:3] == [2.0, 2.0, 2.0]
    np.testing.assert_allclose(sc[3:9], [2 * w[0][3], 2 * w[0][2], 2 * w[0][4], 2 * w[0][2], 2 * w[0][4], 2 * w[0][2] ** 2 * w[0][3]])


def test_cpu_tensors_fail_loudly():
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import _native, loss_functions as lf, test_multiclass, train_multiclass
    x, g = torch.rand(2, 3, 4, 4), torch.rand(2, 3, 4, 4)
    for call in (lambda: eco.losses_fn(x, g), lambda: eco.losses_fn(x, g, True), lambda: lf.dice_loss(x, g),
                 lambda: lf.cross_entropy_loss(x, g), lambda: lf.cross_entropy_loss(x, g, bce=True),
                 lambda: train_multiclass.losses_fn(x, g), lambda: test_multiclass.score_batch(x, g),
                 lambda: eco.binary_cross_entropy(x, g)):
        with pytest.raises(_native.EcoLossError, match="no CPU fallback"):
            call()


def test_reference_quirks_kept_on_the_host_side():
    import ecologysemanticsegmentation_b200 as eco
    from ecologysemanticsegmentation_b200 import loss_functions as lf
    x, g = torch.rand(2, 2, 4, 4), torch.rand(2, 2, 4, 4)
    with pytest.raises(AssertionError, match="Organ ratios size mismatch"):
        eco.losses_fn(x, g, composite_set_theory=True)  # C=2 vs 3 ratios (loss_composite.py:25)
    # the three *_list lambdas are broken in the reference too (loss_functions.py:22-24)
    for fn in (lf.focal_list, lf.classification_dice_list):  # unexpected keyword 'bce'
        with pytest.raises(TypeError):
            fn([x], [g])
    # non-default keywords are accepted (powf instantiation of the kernels) and still need CUDA tensors
    from ecologysemanticsegmentation_b200 import _native
    for call in (lambda: lf.focal_loss(x, g, gamma=2.0), lambda: lf.twersky_loss(x, g, alpha=0.7),
                 lambda: lf.focal_dice_coefficient(x, g, gamma=1.0)):
        with pytest.raises(_native.EcoLossError, match="no CPU fallback"):
            call()


def test_shard_bounds_cover_the_batch():
    from ecologysemanticsegmentation_b200.distributed import shard_bounds
    for n in (1, 7, 54, 432):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_bounds(432, 8, r) for r in range(8)] == [(54 * r, 54 * r + 54) for r in range(8)]
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_synthetic_configs_are_nested_and_seeded():
    from ecologysemanticsegmentation_b200.synthetic import CONFIGS, make_inputs
    z1, g1 = make_inputs(2, 3, 32, 102)
    z2, g2 = make_inputs(2, 3, 32, 102)
    assert torch.equal(z1, z2) and torch.equal(g1, g2)
    assert set(g1.unique().tolist()) <= {0.0, 1.0}
    assert bool((g1[:, 0] >= g1[:, 1]).all()) and bool((g1[:, 1] >= g1[:, 2]).all())
    assert CONFIGS["cfg2"][1:] == (54, 3, 256) and CONFIGS["cfg3"][1:] == (54, 3, 1024)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, call or execute it."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ecologysemanticsegmentation_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle"
                if f.endswith(".py"):
                    assert "torch_port" not in src and "oracle." not in src, f


def test_frame_preprocessing_has_no_cpu_fallback():
    """test_video.preprocess_frames (ess/test_video.py:70-78) is a CUDA kernel: CPU tensors are refused loudly."""
    import torch
    from ecologysemanticsegmentation_b200 import _native as nat
    from ecologysemanticsegmentation_b200 import test_video as tv
    with pytest.raises(nat.EcoLossError):
        tv.preprocess_frames(torch.zeros(8, 8, 3, dtype=torch.uint8), (4, 4))


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    """No library, no result: the product path raises instead of falling back to PyTorch or to the oracle."""
    from ecologysemanticsegmentation_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "libecoloss.so"))
    with pytest.raises(_native.EcoLossError, match="is missing.*no CPU or PyTorch fallback"):
        _native.lib()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (or the C ABI's header) may import or mention it."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ecologysemanticsegmentation_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|importlib.*oracle|__import__\(.oracle")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                for i, line in enumerate(open(os.path.join(dirpath, f), encoding="utf-8"), 1):
                    assert not pat.search(line), f"{f}:{i}: {line.strip()}"
