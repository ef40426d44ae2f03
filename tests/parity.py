"""Shared helpers for the parity tests (definitions from SURVEY.md 8(d) "Parity metric")."""
import numpy as np
import torch

LOSS_NAMES = ["ce", "bce", "focal", "dice", "gdice", "twersky", "focal_dice"]
TOL = 1e-5          # fp32 tolerance stated by BASELINE.json north_star
TOL_BF16 = 1e-2


def as_np(v):
    if isinstance(v, (list, tuple)):
        return np.array([float(t) for t in v], dtype=np.float64)
    if isinstance(v, torch.Tensor):
        return v.detach().double().cpu().numpy()
    return np.asarray(v, dtype=np.float64)


def assert_losses_close(ours, ref, tol=TOL, what=""):
    ours, ref = as_np(ours), as_np(ref)
    assert ours.shape == ref.shape, (ours.shape, ref.shape)
    for k in range(len(ref)):
        if ref[k] == 0.0:
            assert abs(ours[k]) <= 1e-12, f"{what} loss[{k}] expected exactly 0, got {ours[k]}"
        else:
            rel = abs(ours[k] - ref[k]) / abs(ref[k])
            assert rel <= tol, f"{what} loss[{k}] ours={ours[k]!r} ref={ref[k]!r} rel={rel:.3e} > {tol}"


def grad_errors(ours, ref):
    ours, ref = as_np(ours), as_np(ref)
    d = ours - ref
    mx = np.abs(d).max() / max(np.abs(ref).max(), 1e-300)
    l2 = np.sqrt((d * d).sum()) / max(np.sqrt((ref * ref).sum()), 1e-300)
    return mx, l2


def assert_grad_close(ours, ref, tol=TOL, what=""):
    assert tuple(ours.shape) == tuple(ref.shape), (ours.shape, ref.shape)
    mx, l2 = grad_errors(ours, ref)
    assert mx <= tol and l2 <= tol, f"{what} grad max-norm err {mx:.3e}, rel-L2 err {l2:.3e} > {tol}"
    return mx, l2
