"""ctypes binding of ``csrc/libecoloss.so`` (the C ABI declared in ``include/ecoloss.h``).

There is no CPU fallback and no alternative backend: if the extension is missing, or a tensor is
not on a CUDA device, the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libecoloss.so")

ECO_F32, ECO_BF16, ECO_U8 = 0, 1, 2
EVAL_PROBS, EVAL_UNUNION = 1, 2   # flag word of eco_dice_counts (ECO_EVAL_*)
NSTAT, NLOSS, NJAC = 8, 7, 7
C3_NLEAF, C3_NACC = 21, 100
FLAG_A_LOGIT, FLAG_B_LOGIT, FLAG_NEED_BG = 1, 2, 4


class EcoView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sn", C.c_int64), ("sc", C.c_int64), ("dtype", C.c_int32), ("_pad", C.c_int32)]


class EcoOut(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sn", C.c_int64), ("sc", C.c_int64), ("dtype", C.c_int32), ("_pad", C.c_int32)]


class EcoPeerExchange(C.Structure):
    _fields_ = [("peer_xch_dev", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32), ("epoch", C.c_uint32),
                ("_pad", C.c_uint32), ("status", C.c_void_p), ("timeout_ms", C.c_double)]


C3_UNION_LABELS, C3_PROBS, C3_NO_GRAD = 1, 2, 4


class EcoLeafShape(C.Structure):
    """Keyword parameters of the stand-alone primitives (loss_functions.py:46,82,96); defaults = the reference's."""
    _fields_ = [("focal_gamma", C.c_double), ("tversky_alpha", C.c_double), ("tversky_beta", C.c_double),
                ("focal_dice_gamma", C.c_double)]


DEFAULT_SHAPE = (1.5, 0.5, 0.3, 1.8)


class EcoLossError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()

_vp, _i32, _i64, _u32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_double
_VIEW, _OUT, _SHAPE = C.POINTER(EcoView), C.POINTER(EcoOut), C.POINTER(EcoLeafShape)

# name -> (restype, argtypes); must list every symbol of include/ecoloss.h (tests check this)
SIGNATURES = {
    "eco_version": (C.c_char_p, []),
    "eco_last_error": (C.c_char_p, []),
    "eco_sm_count": (C.c_int, [C.c_int]),
    "eco_pair_ws_bytes": (_i64, [_i32]),
    "eco_pair_stats": (C.c_int, [_VIEW, _VIEW, _i32, _i32, _i64, _u32, _vp, _i64, _vp, C.c_int, _vp]),
    "eco_pair_finalize": (C.c_int, [_vp, _i32, _f64, C.POINTER(_f64), _vp, _vp, _vp, C.c_int, _vp]),
    "eco_pair_grad": (C.c_int, [_VIEW, _VIEW, _i32, _i32, _i64, _u32, _vp, _vp, _OUT, _OUT, _i32, C.c_int, _vp]),
    "eco_pair_stats_shaped": (C.c_int, [_VIEW, _VIEW, _i32, _i32, _i64, _u32, _SHAPE, _vp, _i64, _vp, C.c_int, _vp]),
    "eco_pair_finalize_shaped": (C.c_int, [_vp, _i32, _f64, C.POINTER(_f64), _SHAPE, _vp, _vp, _vp, C.c_int, _vp]),
    "eco_pair_grad_shaped": (C.c_int, [_VIEW, _VIEW, _i32, _i32, _i64, _u32, _SHAPE, _vp, _vp, _OUT, _OUT, _i32, C.c_int,
                                       _vp]),
    "eco_pair_fused_ws_bytes": (_i64, [_i32]),
    "eco_pair_fused": (C.c_int, [_VIEW, _VIEW, _i32, _i32, _i64, _u32, _f64, _f64, _SHAPE, _vp, _vp, _i64, _vp, _vp, _OUT, _OUT,
                                 C.c_int, _vp]),
    "eco_pair_fused_ex": (C.c_int, [_VIEW, _VIEW, _i32, _i32, _i64, _u32, _f64, _f64, _SHAPE, _vp, _vp, _vp, _i64, _vp, _vp, _OUT,
                                    _OUT, C.c_int, _vp]),
    "eco_composite3_ws_bytes": (_i64, []),
    "eco_composite3_stats": (C.c_int, [_VIEW, _VIEW, _i32, _i64, _i32, _vp, _i64, _vp, C.c_int, _vp]),
    "eco_composite3_finalize": (C.c_int, [_vp, C.POINTER(_f64), _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "eco_composite3_grad": (C.c_int, [_VIEW, _VIEW, _i32, _i64, _i32, _vp, _vp, _OUT, C.c_int, _vp]),
    "eco_composite3_fused": (C.c_int, [_VIEW, _VIEW, _i32, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _OUT, C.c_int, _vp]),
    "eco_multiclass3_fused": (C.c_int, [_VIEW, _VIEW, _i32, _i64, C.c_double, _vp, _vp, _i64, _vp, _OUT, C.c_int, _vp]),
    "eco_multiclass3_step": (C.c_int, [_VIEW, _VIEW, _i32, _i64, C.c_double, _vp, _vp, _u32, _vp, _i64, _vp, _OUT, C.c_int, _vp]),
    "eco_composite3_fused_sharded": (C.c_int, [_VIEW, _VIEW, _i32, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _OUT, _vp, _i32, _i32,
                                               _u32, C.c_int, _vp]),
    "eco_composite3_step": (C.c_int, [_VIEW, _VIEW, _i32, _i64, _u32, _vp, _vp, _vp, _i64, _vp, _OUT, C.POINTER(EcoPeerExchange),
                                      C.c_int, _vp]),
    "eco_composite3_step_if_changed": (C.c_int, [_VIEW, _VIEW, _i32, _i64, _u32, _vp, _vp, _vp, _vp, _i64, _vp, _OUT, C.c_int, _vp]),
    "eco_xch_poll_status": (C.c_int, [_vp, _i64, C.POINTER(C.c_uint32), C.c_int, _vp]),
    "eco_xch_bytes": (_i64, [_i32]),
    "eco_xch_alloc": (C.c_int, [_i32, C.POINTER(_vp), C.c_char_p, C.c_int]),
    "eco_xch_open": (C.c_int, [C.c_char_p, C.POINTER(_vp), C.c_int]),
    "eco_xch_close": (C.c_int, [_vp, C.c_int]),
    "eco_xch_free": (C.c_int, [_vp, C.c_int]),
    "eco_dice_ws_bytes": (_i64, [_i32, _i32]),
    "eco_dice_counts": (C.c_int, [_VIEW, _VIEW, _i32, _i32, _i64, _vp, _i32, _i32, _vp, _i64, _vp, _vp, C.c_int, _vp]),
    "eco_dice_counts_ex": (C.c_int, [_VIEW, _VIEW, _i32, _i32, _i64, _vp, _i32, _i32, _vp, _i64, _vp, _vp, _vp, C.c_int, _vp]),
    "eco_dice_finalize_ex": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, C.c_int, _vp]),
    "eco_dice_finalize": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, C.c_int, _vp]),
    "eco_masks_u8": (C.c_int, [_VIEW, _i32, _i32, _i64, C.c_float, _i32, _i32, _vp, C.c_int, _vp]),
    "eco_frames_plan_sizes": (C.c_int, [_i32, _i32, _i32, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "eco_frames_plan": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i32), C.POINTER(_i32)]),
    "eco_frames_preprocess": (C.c_int, [_vp, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32,
                                        _vp, _vp, C.c_int, _vp]),
    "eco_union_sets": (C.c_int, [_vp, _i32, _i64, _i32, _i64, _i64, _i64, C.c_uint64, _i32, C.c_int, _vp]),
    "eco_softce_ws_bytes": (_i64, []),
    "eco_softce_stats": (C.c_int, [_VIEW, _VIEW, _i32, _i32, _i64, _i32, _vp, _i64, _vp, C.c_int, _vp]),
    "eco_softce_grad": (C.c_int, [_VIEW, _VIEW, _i32, _i32, _i64, _f64, _f64, _vp, _OUT, _OUT, C.c_int, _vp]),
}


def lib():
    """The loaded library; raises if it has not been built (``python -m ecologysemanticsegmentation_b200.csrc.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise EcoLossError(
                    f"{LIB_PATH} is missing: build the CUDA extension first "
                    "(python -m ecologysemanticsegmentation_b200.csrc.build). There is no CPU or PyTorch fallback.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)  # AttributeError here = header/library mismatch: fail loudly
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().eco_last_error().decode("utf-8", "replace")
        raise EcoLossError(f"{what} failed (rc={rc}): {msg}")


def dtype_code(t: torch.Tensor, allow_u8: bool = False) -> int:
    if t.dtype == torch.float32:
        return ECO_F32
    if t.dtype == torch.bfloat16:
        return ECO_BF16
    if allow_u8 and t.dtype == torch.uint8:
        return ECO_U8
    raise EcoLossError(f"unsupported dtype {t.dtype}: the kernels take float32 or bfloat16")


def require_cuda(*tensors):
    for t in tensors:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"expected a torch.Tensor, got {type(t)}")
        if not t.is_cuda:
            raise EcoLossError(
                "ecologysemanticsegmentation_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU fallback"
                % t.device.type)


def current_stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def planes(t: torch.Tensor):
    """Describe a 4-D tensor [N,C,H,W] whose (H,W) planes are contiguous as (tensor, sn, sc) in elements,
    copying only when the planes themselves are strided."""
    assert t.dim() == 4
    n, c, h, w = t.shape
    st = t.stride()
    ok = (w == 1 or st[3] == 1) and (h == 1 or st[2] == w)
    if not ok or (n > 1 and st[0] < 0) or (c > 1 and st[1] < 0):
        t = t.contiguous()
        st = t.stride()
    return t, (st[0] if n > 1 else c * h * w), (st[1] if c > 1 else h * w)


def view_of(t: torch.Tensor, sn: int, sc: int, allow_u8: bool = False) -> EcoView:
    return EcoView(t.data_ptr(), sn, sc, dtype_code(t, allow_u8), 0)


def out_of(t, sn: int = 0, sc: int = 0) -> EcoOut:
    if t is None:
        return EcoOut(None, 0, 0, 0, 0)
    return EcoOut(t.data_ptr(), sn, sc, dtype_code(t), 0)


# ------------------------------------------------------------------------------------------------
# workspaces: zero-initialised once, re-armed by the kernels themselves; one per (kind, device, stream)
# ------------------------------------------------------------------------------------------------
_workspaces = {}


def workspace(kind: str, nbytes: int, device: torch.device) -> torch.Tensor:
    key = (kind, device.index, current_stream_ptr(device))
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def reset_workspaces():
    """Drop cached workspaces (call after a failed launch left arrival counters armed)."""
    _workspaces.clear()
