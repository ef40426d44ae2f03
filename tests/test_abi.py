"""The C-ABI library loads without a GPU and exports every symbol include/ecoloss.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ecoloss.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eco_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_and_exports_header_symbols():
    from ecologysemanticsegmentation_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        from ecologysemanticsegmentation_b200.csrc import build
        build.build()
    handle = ctypes.CDLL(_native.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/ecoloss.h but not exported"
    # the Python binding table covers exactly the header
    assert sorted(_native.SIGNATURES) == syms


def test_version_and_error_string_without_gpu():
    from ecologysemanticsegmentation_b200 import _native
    L = _native.lib()
    assert b"sm_100a" in L.eco_version()
    assert L.eco_pair_ws_bytes(3) > 0 and L.eco_pair_ws_bytes(0) < 0
    assert L.eco_composite3_ws_bytes() > 0
    assert L.eco_dice_ws_bytes(3, 19) > 0 and L.eco_dice_ws_bytes(3, 21) < 0
    assert L.eco_softce_ws_bytes() > 0


def test_argument_errors_are_reported_not_crashed():
    """Argument validation happens before any CUDA call, so it can be exercised on a CPU-only box."""
    from ecologysemanticsegmentation_b200 import _native
    L = _native.lib()
    rc = L.eco_pair_stats(None, None, 1, 1, 16, 0, None, 0, None, 0, None)
    assert rc < 0 and b"null" in L.eco_last_error()
    v = _native.EcoView(1 << 20, 16, 16, 0, 0)
    rc = L.eco_pair_stats(ctypes.byref(v), ctypes.byref(v), 0, 1, 16, 0, None, 0, None, 0, None)
    assert rc < 0 and b"empty" in L.eco_last_error()
    rc = L.eco_dice_counts(ctypes.byref(v), ctypes.byref(v), 1, 1, 16, None, 25, 0, None, 0, None, None, 0, None)
    assert rc < 0 and b"n_thr" in L.eco_last_error()
    # flag word of the scoring call: unknown bits, and the fused prediction un-union together with thresholds
    thr = (ctypes.c_float * 1)(0.8)
    rc = L.eco_dice_counts(ctypes.byref(v), ctypes.byref(v), 1, 3, 16, None, 0, 8, None, 0, None, None, 0, None)
    assert rc < 0 and b"flag" in L.eco_last_error()
    rc = L.eco_dice_counts(ctypes.byref(v), ctypes.byref(v), 1, 3, 16, thr, 1, _native.EVAL_UNUNION, None, 0, None, None, 0, None)
    assert rc < 0 and b"un-union" in L.eco_last_error()


def test_new_entry_points_validate_arguments_without_gpu():
    from ecologysemanticsegmentation_b200 import _native
    L = _native.lib()
    assert L.eco_pair_fused_ws_bytes(1) > L.eco_pair_ws_bytes(1) and L.eco_pair_fused_ws_bytes(65) < 0
    v = _native.EcoView(1 << 20, 16, 16, 0, 0)
    rc = L.eco_pair_fused(ctypes.byref(v), ctypes.byref(v), 1, 65, 16, 0, 0.0, 1.0, None, None, None, 0, None, None, None, None,
                          0, None)
    assert rc < 0 and b"C must be <= 64" in L.eco_last_error()
    rc = L.eco_pair_fused(ctypes.byref(v), ctypes.byref(v), 1, 1, 16, 0, 0.0, 1.0, None, None, None, 0, None, None, None, None,
                          0, None)
    assert rc < 0 and b"null" in L.eco_last_error()
    rc = L.eco_masks_u8(None, 1, 1, 16, 0.0, 0, 0, None, 0, None)
    assert rc < 0 and b"null" in L.eco_last_error()
    rc = L.eco_masks_u8(ctypes.byref(v), 0, 1, 16, 0.0, 0, 0, None, 0, None)
    assert rc < 0 and b"empty" in L.eco_last_error()
    shape = _native.EcoLeafShape(2.0, 0.7, 0.3, 1.0)
    rc = L.eco_pair_stats_shaped(None, None, 1, 1, 16, 0, ctypes.byref(shape), None, 0, None, 0, None)
    assert rc < 0 and b"null" in L.eco_last_error()
