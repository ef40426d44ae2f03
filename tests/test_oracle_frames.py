"""The frame pre-processing oracle (oracle/frames.py: Pillow's 8-bit bilinear resampling + torchvision's ToTensor /
Normalize, ess/test_video.py:70-78) against the committed golden vectors, against Pillow + torchvision themselves where
they are installed, and the host-side plan of the C library (eco_frames_plan: no GPU involved) against the oracle's tables."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from oracle import frames as of

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_frames.npz"))
SMALL = {11: (32, 32), 12: (64, 64), 13: (40, 56), 14: (64, 64), 15: (7, 5)}


def _golden_frame(seed, h, w):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_frames", os.path.join(os.path.dirname(__file__), "golden", "make_golden_frames.py"))
    # only the seeded input generator is needed; Pillow / torchvision are imported by that module
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.frame(seed, h, w)


@pytest.mark.parametrize("seed", sorted(SMALL))
def test_oracle_equals_golden_small(seed):
    got = of.preprocess(GOLD[f"in_{seed}"], SMALL[seed])
    ref = GOLD[f"out_{seed}"]
    assert got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_oracle_equals_pillow_and_torchvision_when_installed():
    Image = pytest.importorskip("PIL.Image")
    transforms = pytest.importorskip("torchvision.transforms")
    rs = np.random.RandomState(5)
    for (h, w), size in [((270, 480), (128, 128)), ((33, 71), (50, 20)), ((128, 128), (256, 256)), ((256, 300), (256, 256))]:
        a = (rs.rand(h, w, 3) * 255).astype(np.uint8)
        im = Image.fromarray(a)
        assert np.array_equal(np.asarray(im.resize((size[1], size[0]), Image.BILINEAR)), of.resize_bilinear_u8(a, size))
        tf = transforms.Compose([transforms.Resize(size), transforms.ToTensor(),
                                 transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        assert np.array_equal(tf(im).numpy().view(np.uint32), of.preprocess(a, size).view(np.uint32))


def test_oracle_equals_golden_digest_1080p():
    pytest.importorskip("PIL.Image")
    pytest.importorskip("torchvision.transforms")
    for item in GOLD["big"]:
        seed, size, digest = str(item).split(":")
        a = _golden_frame(int(seed), 1080, 1920)
        got = of.preprocess(a, (int(size), int(size)))
        assert hashlib.sha256(got.tobytes()).hexdigest() == digest


@pytest.mark.parametrize("sizes", [(1080, 1920, 512, 512), (1080, 1920, 256, 256), (37, 53, 256, 256), (300, 200, 64, 96),
                                   (256, 256, 256, 256), (19, 23, 7, 5)])
def test_library_plan_equals_oracle_tables(sizes):
    from ecologysemanticsegmentation_b200 import _native as nat
    hin, win, hout, wout = sizes
    L = nat.lib()
    ksx, ksy = C.c_int32(), C.c_int32()
    assert L.eco_frames_plan_sizes(hin, win, hout, wout, C.byref(ksx), C.byref(ksy)) == 0
    xb, kx = np.zeros((wout, 2), np.int32), np.zeros((wout, ksx.value), np.int32)
    yb, ky = np.zeros((hout, 2), np.int32), np.zeros((hout, ksy.value), np.int32)
    lut = np.zeros((3, 256), np.float32)
    m, s = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)
    pc, pr = C.c_int32(), C.c_int32()
    assert L.eco_frames_plan(hin, win, hout, wout, m.ctypes.data, s.ctypes.data, xb.ctypes.data, kx.ctypes.data, yb.ctypes.data,
                             ky.ctypes.data, lut.ctypes.data, C.byref(pc), C.byref(pr)) == 0
    bx, kkx = of.precompute_coeffs(win, wout)
    by, kky = of.precompute_coeffs(hin, hout)
    assert np.array_equal(xb, bx) and np.array_equal(kx, kkx) and np.array_equal(yb, by) and np.array_equal(ky, kky)
    ref = of.to_tensor_normalize(np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(3, 2), m, s)[:, 0, :]
    assert np.array_equal(lut.view(np.uint32), ref.view(np.uint32))
    assert pc.value >= 1 and pr.value >= 1
