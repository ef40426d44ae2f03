"""TEST INFRASTRUCTURE (oracle): CPU restatement of the frame pre-processing of ess/test_video.py:70-78

    transforms.Resize((256, 256)) -> transforms.ToTensor() -> transforms.Normalize(mean, std)

on a PIL RGB image.  The arithmetic lives in third-party code that is not part of /root/reference:
  * Pillow (Image.resize with BILINEAR; src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
    ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc) -- a separable triangle filter whose support grows with
    the down-scaling factor, coefficients normalised in double, rounded to 22-bit fixed point, one rounding to uint8 after
    the horizontal pass and one after the vertical pass;
  * torchvision (F.to_tensor: uint8 HWC -> float32 CHW / 255; F.normalize: (x - mean) / std in float32).
Pinned against Pillow + torchvision themselves in tests/test_oracle_frames.py (here: Pillow 12.2.0, torchvision 0.26.0) and
against tests/golden/golden_frames.npz.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c:precompute_coeffs for the whole axis (box = (0, in_size)), then normalize_coeffs_8bpc.
    Returns (bounds int32 [out, 2] = (first input index, tap count), coefficients int32 [out, ksize])."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bilinear((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One 8-bit pass along `axis` of an [H, W, C] uint8 image: ss = 2^21 + sum pixel * k; clip8(ss >> 22)."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for o in range(bounds.shape[0]):
        lo, cnt = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.tensordot(kk[o, :cnt].astype(np.int64), src[lo:lo + cnt], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img, out_hw):
    """PIL Image.resize((W, H), BILINEAR) of an [H, W, C] uint8 array: horizontal pass first, then vertical (a pass whose
    size does not change is skipped, as ImagingResample does)."""
    h, w = img.shape[:2]
    oh, ow = out_hw
    if ow != w:
        bx, kx = precompute_coeffs(w, ow)
        img = _pass(img, bx, kx, 1)
    if oh != h:
        by, ky = precompute_coeffs(h, oh)
        img = _pass(img, by, ky, 0)
    return img


def to_tensor_normalize(img_u8, mean, std):
    """torchvision F.to_tensor + F.normalize on an [H, W, C] uint8 array: float32 [C, H, W]; x / 255, then (x - mean) / std,
    every operation rounded to float32."""
    x = np.ascontiguousarray(np.moveaxis(img_u8, 2, 0)).astype(np.float32) / np.float32(255.0)
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    return ((x - m) / s).astype(np.float32)


def preprocess(img_u8, out_hw=(256, 256), mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """ess/test_video.py:70-78 without the file I/O and the batch dimension."""
    return to_tensor_normalize(resize_bilinear_u8(img_u8, out_hw), mean, std)
