"""float64 sufficient-statistics formulation of the loss path (numpy).  TEST INFRASTRUCTURE.

This is the algorithm the CUDA kernels implement, restated on the CPU so it can be checked
against ``oracle.torch_port`` (= the reference, bit for bit) independently of any GPU:

    per (a, b) leaf:   sums  = [n, Sa, Sb, Sab, Sbb, SP, FL, FLB]            (one streaming pass)
                       losses = closed forms of the sums                     
                       grads  = affine in (a, b) + sigmoid/focal' terms with global coefficients

Reference lines each closed form follows are cited on the function.  Also used by the world-size-2
``gloo`` tests: per-shard sums are additive, so shard -> sum -> allreduce -> ``leaf_losses`` must equal
the full-batch value.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-7
M_DICE = 10 * 0.33          # classification_dice_loss(factor=10): ess/loss_functions.py:116
ALPHA, BETA = 0.5, 0.3      # twersky_loss defaults: ess/loss_functions.py:82
G_FOCAL, G_FDICE = 1.5, 1.8  # focal_loss / focal_dice_coefficient gammas: :46, :96

# index of each statistic in a leaf's sums vector (shared with csrc/eco_common.cuh)
N_, SA, SB, SAB, SBB, SP, FL, FLB = range(8)
NSTAT = 8


def _f64(t):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float64)


def softplus_term(b):
    """max(b,0) + log(1+exp(-|b|)): BCEWithLogits' input-only part (ess/__init__.py:24)."""
    return np.maximum(b, 0) + np.log1p(np.exp(-np.abs(b)))


def focal_fg(b):
    """-(1-b)^1.5 log(b+eps): ess/loss_functions.py:47."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return -np.power(1 - b, G_FOCAL) * np.log(b + EPS)


def focal_bg(b):
    """-b^1.5 log(1-b+eps): ess/loss_functions.py:48 (without the background weight)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return -np.power(b, G_FOCAL) * np.log(1 - b + EPS)


def leaf_sums(a, b):
    a, b = _f64(a).ravel(), _f64(b).ravel()
    s = np.zeros(NSTAT)
    s[N_] = a.size
    s[SA], s[SB], s[SAB], s[SBB] = a.sum(), b.sum(), (a * b).sum(), (b * b).sum()
    s[SP], s[FL], s[FLB] = softplus_term(b).sum(), focal_fg(b).sum(), focal_bg(b).sum()
    return s


def _phi(t):
    return -np.power(1 - t, G_FDICE) * np.log(t + EPS)


def _dphi(t):
    return G_FDICE * np.power(1 - t, G_FDICE - 1) * np.log(t + EPS) - np.power(1 - t, G_FDICE) / (t + EPS)


def _moments(s):
    n = s[N_]
    I, D = s[SAB], s[SA] + s[SBB]
    Ib = n - s[SA] - s[SB] + s[SAB]
    Db = (n - s[SA]) + (n - 2 * s[SB] + s[SBB])
    FN, FP = s[SA] - s[SAB], s[SB] - s[SAB]
    return n, I, D, Ib, Db, FN, FP


def leaf_losses(s, bw=0.0, scale=1.0):
    """[ce, bce, focal, dice, gdice, tversky, focal_dice] * scale from one leaf's sums.
    ess/loss_composite.py:32-39 with loss_functions.py:26-117; CE is identically 0 on 1-channel slices."""
    n, I, D, Ib, Db, FN, FP = _moments(s)
    bce = (s[SP] - s[SAB]) / n
    fl = (s[FL] + bw * s[FLB]) / n
    dice = -(2 * I + EPS) / (D + EPS) - bw * (2 * Ib + EPS) / (2 * Db + EPS)
    gdice = -((I + EPS) / (D + EPS) + bw * (Ib + EPS) / (Db + EPS))
    t1 = I + ALPHA * FN + BETA * FP + EPS
    t2 = Ib + ALPHA * FP + BETA * FN + EPS
    tv = -(I + EPS) / t1 + bw * (-(Ib + EPS) / t2)
    dc, dcb = (2 * I + EPS) / (D + EPS), (2 * Ib + EPS) / (Db + EPS)
    fd = _phi(dc) + bw * _phi(dcb)
    return scale * np.array([0.0, bce, fl, M_DICE * dice, M_DICE * gdice, M_DICE * tv, M_DICE * fd])


def leaf_coefs(s, w, bw=0.0):
    """d(sum_k w_k L_k)/d(stat) for stats [Sa, Sb, Sab, Sbb, SP, FL, FLB] (w already includes the scale)."""
    n, I, D, Ib, Db, FN, FP = _moments(s)
    m = M_DICE
    dI = dD = dIb = dDb = dFN = dFP = 0.0
    # dice (loss_functions.py:55-63)
    dI += -m * w[3] * 2 / (D + EPS)
    dD += m * w[3] * (2 * I + EPS) / (D + EPS) ** 2
    dIb += -m * w[3] * bw * 2 / (2 * Db + EPS)
    dDb += m * w[3] * bw * (2 * Ib + EPS) * 2 / (2 * Db + EPS) ** 2
    # generalized dice (:69-80)
    dI += -m * w[4] / (D + EPS)
    dD += m * w[4] * (I + EPS) / (D + EPS) ** 2
    dIb += -m * w[4] * bw / (Db + EPS)
    dDb += m * w[4] * bw * (Ib + EPS) / (Db + EPS) ** 2
    # tversky (:84-94)
    t1 = I + ALPHA * FN + BETA * FP + EPS
    t2 = Ib + ALPHA * FP + BETA * FN + EPS
    dI += m * w[5] * (-1 / t1 + (I + EPS) / t1 ** 2)
    dIb += m * w[5] * bw * (-1 / t2 + (Ib + EPS) / t2 ** 2)
    dFN += m * w[5] * ((I + EPS) * ALPHA / t1 ** 2 + bw * (Ib + EPS) * BETA / t2 ** 2)
    dFP += m * w[5] * ((I + EPS) * BETA / t1 ** 2 + bw * (Ib + EPS) * ALPHA / t2 ** 2)
    # focal dice (:98-108)
    dc, dcb = (2 * I + EPS) / (D + EPS), (2 * Ib + EPS) / (Db + EPS)
    p1 = _dphi(dc)
    dI += m * w[6] * p1 * 2 / (D + EPS)
    dD += -m * w[6] * p1 * (2 * I + EPS) / (D + EPS) ** 2
    if bw != 0:
        p2 = _dphi(dcb)
        dIb += m * w[6] * bw * p2 * 2 / (Db + EPS)
        dDb += -m * w[6] * bw * p2 * (2 * Ib + EPS) / (Db + EPS) ** 2
    c_sa = dD - dIb - dDb + dFN
    c_sb = -dIb - 2 * dDb + dFP
    c_sab = dI + dIb - dFN - dFP - w[1] / n          # BCE's -b*a term (ess/__init__.py:24)
    c_sbb = dD + dDb
    return np.array([c_sa, c_sb, c_sab, c_sbb, w[1] / n, w[2] / n, w[2] * bw / n])


def dfocal_fg(b):
    with np.errstate(invalid="ignore", divide="ignore"):
        return G_FOCAL * np.power(1 - b, G_FOCAL - 1) * np.log(b + EPS) - np.power(1 - b, G_FOCAL) / (b + EPS)


def dfocal_bg(b):
    with np.errstate(invalid="ignore", divide="ignore"):
        return -G_FOCAL * np.power(b, G_FOCAL - 1) * np.log(1 - b + EPS) + np.power(b, G_FOCAL) / (1 - b + EPS)


def sigmoid(b):
    return 1.0 / (1.0 + np.exp(-b))


def leaf_grad(a, b, c):
    """Per-element (dT/da, dT/db) from the coefficient vector of ``leaf_coefs``."""
    a, b = _f64(a), _f64(b)
    ga = c[0] + c[2] * b
    gb = c[1] + c[2] * a + 2 * c[3] * b + c[4] * sigmoid(b)
    if c[5] != 0:
        gb = gb + c[5] * dfocal_fg(b)
    if c[6] != 0:
        gb = gb + c[6] * dfocal_bg(b)
    return ga, gb


# ----------------------------------------------------------------------------------------------
# composite (ess/loss_composite.py:21-94)
# ----------------------------------------------------------------------------------------------
def union_operand(sp, p):
    return sp * (1 - p) + (sp * p + p) * 0.5


def pair_weights(ratios, early_stopped=False):
    """Host-side weight draw, RNG-stream compatible with ess/loss_composite.py:49-52.
    Returns {(i,j): (w_i, w_j, w_d)}."""
    out = {}
    C = len(ratios)
    es = int(early_stopped)
    for i in range(C - 1):
        for j in range(i + 1, C):
            w_i = (1 / ratios[i]) * (1 - es * np.random.choice([0, 1]) * np.random.rand())
            w_j = (1 / ratios[j]) * (1 - es * np.random.choice([0, 1]) * np.random.rand())
            w_d = (1 / (ratios[i] - ratios[j])) * (1 - es * np.random.choice([0, 1]) * np.random.rand())
            out[(i, j)] = (w_i, w_j, w_d)
    return out


def composite_leaves(x, g, weights, scale=2.0):
    """Enumerate every leaf of losses_fn(x, g, composite_set_theory=True) as
    (a, b, leaf_scale, backprop) where ``backprop(ga, gb)`` returns {channel: dT/dx_channel}.
    x, g: float64 arrays [N,C,H,W]."""
    C = x.shape[1]
    leaves = []
    for c in range(C):  # :28 -- a = label, b = prediction
        leaves.append((g[:, c], x[:, c], scale, (lambda ga, gb, c=c: {c: gb})))
    for (i, j), (w_i, w_j, w_d) in weights.items():
        xi, xj, gi, gj = x[:, i], x[:, j], g[:, i], g[:, j]
        s = np.sign(xi - xj)
        d = np.abs(xi - xj)
        gd = np.abs(gi - gj)
        q = d * xi
        # I1 (:56)
        leaves.append((xi * xj, gj, scale * w_j, (lambda ga, gb, i=i, j=j, xi=xi, xj=xj: {i: ga * xj, j: ga * xi})))
        # U1 (:59)
        leaves.append((gi, union_operand(xi, xj), scale * w_i,
                       (lambda ga, gb, i=i, j=j, xi=xi, xj=xj: {i: gb * (1 - 0.5 * xj), j: gb * 0.5 * (1 - xi)})))
        # I2 (:63-65)
        leaves.append((xi * d, gd, scale * w_d,
                       (lambda ga, gb, i=i, j=j, xi=xi, d=d, s=s: {i: ga * (d + xi * s), j: ga * (-xi * s)})))
        # U2 (:68-70)
        leaves.append((gi, union_operand(xi, d), scale * w_i,
                       (lambda ga, gb, i=i, j=j, xi=xi, d=d, s=s:
                        {i: gb * ((1 - 0.5 * d) + 0.5 * (1 - xi) * s), j: gb * (0.5 * (1 - xi) * (-s))})))
        # I3 (:74-76)
        leaves.append((xi * q, gd, scale * w_d,
                       (lambda ga, gb, i=i, j=j, xi=xi, d=d, s=s:
                        {i: ga * (2 * xi * d + xi * xi * s), j: ga * (-xi * xi * s)})))
        # U3 (:79-81)
        leaves.append((gi, union_operand(xi, q), scale * w_i * w_i * w_j,
                       (lambda ga, gb, i=i, j=j, xi=xi, d=d, s=s, q=q:
                        {i: gb * ((1 - 0.5 * q) + 0.5 * (1 - xi) * (d + xi * s)),
                         j: gb * (0.5 * (1 - xi) * (-xi * s))})))
    return leaves


def composite_losses_and_grad(x, g, weights, upstream=None, scale=2.0):
    """losses[7] of the composite loss and, if ``upstream`` (7 weights) is given, dT/dx."""
    x, g = _f64(x), _f64(g)
    total = np.zeros(7)
    gx = np.zeros_like(x) if upstream is not None else None
    for a, b, sc, back in composite_leaves(x, g, weights, scale):
        s = leaf_sums(a, b)
        total += leaf_losses(s, 0.0, sc)
        if upstream is not None:
            c = leaf_coefs(s, sc * np.asarray(upstream, dtype=np.float64), 0.0)
            ga, gb = leaf_grad(a, b, c)
            for ch, v in back(ga, gb).items():
                gx[:, ch] += v
    return total, gx


def plain_losses_and_grad(x, g, upstream=None, scale=2.0):
    """C>1 without composite terms: sum over channels of leaf(a=g_c, b=x_c)."""
    x, g = _f64(x), _f64(g)
    total = np.zeros(7)
    gx = np.zeros_like(x) if upstream is not None else None
    for c in range(x.shape[1]):
        s = leaf_sums(g[:, c], x[:, c])
        total += leaf_losses(s, 0.0, scale)
        if upstream is not None:
            co = leaf_coefs(s, scale * np.asarray(upstream, dtype=np.float64), 0.0)
            gx[:, c] = leaf_grad(g[:, c], x[:, c], co)[1]
    return total, gx
