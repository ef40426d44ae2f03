"""Frame pre- and post-processing around the network of ess/test_video.py, on the GPU.

    ess/test_video.py:70-78    preprocess_image: Resize((256, 256)) -> ToTensor() -> Normalize(mean, std) on a PIL image
    ess/test_video.py:129-130  (output_image * 255).astype(np.uint8)

The reference resizes with Pillow on the host and copies float32 tensors (4 B/element) to the GPU; here the uint8 frames
(3 B/pixel) are copied as they are and ONE kernel (csrc/eco_frames.cu) does Pillow's two 8-bit resampling passes, the
division by 255 and the normalisation, bit-identical to Pillow + torchvision on the CPU.  No CPU fallback: the frames must
be CUDA tensors (``preprocess_image`` reads the file with Pillow and moves the bytes to the current CUDA device)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nat
from .test_multiclass import to_uint8_masks

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # test_video.py:74
IMAGENET_STD = (0.229, 0.224, 0.225)

_plans = {}


class _Plan:
    """Pillow's coefficient tables for one (input size, output size, mean, std), built on the host by eco_frames_plan (C
    double arithmetic in Pillow's operation order) and kept on the device."""

    def __init__(self, hin, win, hout, wout, mean, std, device):
        L = nat.lib()
        ksx, ksy = C.c_int32(), C.c_int32()
        nat.check(L.eco_frames_plan_sizes(hin, win, hout, wout, C.byref(ksx), C.byref(ksy)), "eco_frames_plan_sizes")
        self.ksx, self.ksy = ksx.value, ksy.value
        xb = np.zeros((wout, 2), np.int32)
        kx = np.zeros((wout, self.ksx), np.int32)
        yb = np.zeros((hout, 2), np.int32)
        ky = np.zeros((hout, self.ksy), np.int32)
        lut = np.zeros((3, 256), np.float32)
        m = np.asarray(mean, np.float32)
        s = np.asarray(std, np.float32)
        pc, pr = C.c_int32(), C.c_int32()
        nat.check(L.eco_frames_plan(hin, win, hout, wout, m.ctypes.data, s.ctypes.data, xb.ctypes.data, kx.ctypes.data,
                                    yb.ctypes.data, ky.ctypes.data, lut.ctypes.data, C.byref(pc), C.byref(pr)), "eco_frames_plan")
        self.patch_cols, self.patch_rows = pc.value, pr.value
        self.host = (xb, kx, yb, ky, lut)
        self.dev = tuple(torch.from_numpy(a).to(device) for a in self.host)


def _plan(hin, win, hout, wout, mean, std, device):
    key = (hin, win, hout, wout, tuple(float(v) for v in mean), tuple(float(v) for v in std), device.index)
    p = _plans.get(key)
    if p is None:
        if len(_plans) >= 32:
            _plans.clear()
        p = _plans[key] = _Plan(hin, win, hout, wout, mean, std, device)
    return p


def preprocess_frames(frames, size=(256, 256), mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """uint8 RGB frames [N, H, W, 3] (or one frame [H, W, 3]) on a CUDA device -> float32 [N, 3, size[0], size[1]]:
    the transform of ess/test_video.py:71-75 applied to every frame."""
    if not isinstance(frames, torch.Tensor) or not frames.is_cuda:
        raise nat.EcoLossError("preprocess_frames expects a CUDA uint8 tensor (no CPU fallback)")
    if frames.dtype != torch.uint8:
        raise TypeError(f"frames must be uint8, got {frames.dtype}")
    if frames.dim() == 3:
        frames = frames.unsqueeze(0)
    if frames.dim() != 4 or frames.shape[3] != 3:
        raise ValueError(f"frames must be [N, H, W, 3] RGB, got {tuple(frames.shape)}")
    if frames.stride(3) != 1 or frames.stride(2) != 3:
        frames = frames.contiguous()
    n, hin, win, _ = frames.shape
    hout, wout = int(size[0]), int(size[1])
    dev = frames.device
    p = _plan(hin, win, hout, wout, mean, std, dev)
    if out is None:
        out = torch.empty((n, 3, hout, wout), dtype=torch.float32, device=dev)
    elif out.shape != (n, 3, hout, wout) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
        raise ValueError("out must be a contiguous float32 [N, 3, H, W] tensor on the frames' device")
    xb, kx, yb, ky, lut = p.dev
    rc = nat.lib().eco_frames_preprocess(frames.data_ptr(), n, hin, win, frames.stride(0), frames.stride(1), xb.data_ptr(),
                                         kx.data_ptr(), p.ksx, yb.data_ptr(), ky.data_ptr(), p.ksy, hout, wout, p.patch_cols,
                                         p.patch_rows, lut.data_ptr(), out.data_ptr(), dev.index,
                                         torch._C._cuda_getCurrentRawStream(dev.index))
    nat.check(rc, "eco_frames_preprocess")
    return out


def preprocess_image(image_path, size=(256, 256)):
    """ess/test_video.py:70-78: the image file -> normalised float32 [1, 3, 256, 256] (on the current CUDA device)."""
    from PIL import Image   # file decoding stays on the host, as in the reference (:76)
    image = Image.open(image_path)
    if image.mode != "RGB":
        image = image.convert("RGB")
    frame = torch.from_numpy(np.asarray(image).copy()).cuda(non_blocking=True)
    return preprocess_frames(frame, size)


def to_uint8_image(output):
    """ess/test_video.py:129-130: ``(output.squeeze().numpy() * 255).astype(np.uint8)`` on the device (1 B/element to copy
    back instead of 4)."""
    x = output
    while x.dim() < 4:
        x = x.unsqueeze(0)
    return to_uint8_masks(x, None, inputs_are_probs=True).squeeze()
