"""Frame pre-processing kernel (csrc/eco_frames.cu; ess/test_video.py:70-78) against the oracle (Pillow's resampling +
torchvision's ToTensor / Normalize restated in oracle/frames.py, pinned against both libraries) and the golden vectors:
every output float bit-identical."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_frames.npz"))
SMALL = {11: (32, 32), 12: (64, 64), 13: (40, 56), 14: (64, 64), 15: (7, 5)}


def _bits(t):
    return t.detach().cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("seed", sorted(SMALL))
def test_frames_equal_golden(seed):
    from ecologysemanticsegmentation_b200 import test_video as tv
    out = tv.preprocess_frames(torch.from_numpy(GOLD[f"in_{seed}"]).cuda(), SMALL[seed])
    assert out.shape == (1, 3) + SMALL[seed]
    assert np.array_equal(_bits(out[0]), GOLD[f"out_{seed}"].view(np.uint32))


@pytest.mark.parametrize("case", [((1080, 1920), (512, 512), 2), ((1080, 1920), (256, 256), 1), ((720, 1280), (512, 512), 3),
                                  ((37, 53), (256, 256), 2), ((301, 203), (65, 97), 2), ((256, 256), (256, 256), 1),
                                  ((512, 700), (512, 512), 1), ((64, 48), (1, 1), 1), ((5, 7), (33, 17), 2)])
def test_frames_equal_oracle(case):
    """down- and up-scaling, ragged tiles, batches, sizes whose rows are not 4-byte aligned"""
    from ecologysemanticsegmentation_b200 import test_video as tv
    from oracle import frames as of
    (h, w), size, n = case
    rs = np.random.RandomState(h * 31 + w)
    a = (rs.rand(n, h, w, 3) * 255).astype(np.uint8)
    a[:, : max(1, h // 8)] = 255
    a[:, :, w - max(1, w // 8):] = 0
    out = tv.preprocess_frames(torch.from_numpy(a).cuda(), size)
    for i in range(n):
        assert np.array_equal(_bits(out[i]), of.preprocess(a[i], size).view(np.uint32)), f"frame {i} of {case}"


def test_frames_strided_and_unaligned_inputs():
    """frames cut out of a larger buffer: row / frame strides larger than the data and a base pointer that is not 4-byte
    aligned; the kernel reads whole 4-byte words but never outside the caller's buffer (first and last word byte by byte)"""
    from ecologysemanticsegmentation_b200 import test_video as tv
    from oracle import frames as of
    rs = np.random.RandomState(3)
    big = (rs.rand(3, 130, 171, 3) * 255).astype(np.uint8)
    dev = torch.from_numpy(big).cuda()
    crop = dev[:, 7:118, 5:166, :]            # [3, 111, 161, 3]: base offset 7*171*3 + 15 bytes, row stride 513 bytes
    out = tv.preprocess_frames(crop, (48, 80))
    for i in range(3):
        assert np.array_equal(_bits(out[i]), of.preprocess(big[i, 7:118, 5:166], (48, 80)).view(np.uint32))
    flat = torch.from_numpy(np.concatenate([np.zeros(1, np.uint8), big[0].reshape(-1)])).cuda()
    odd = flat[1:].view(130, 171, 3)          # the whole allocation, shifted by one byte
    out = tv.preprocess_frames(odd, (40, 40))
    assert np.array_equal(_bits(out[0]), of.preprocess(big[0], (40, 40)).view(np.uint32))


def test_frames_1080p_digest_and_image_file(tmp_path):
    from ecologysemanticsegmentation_b200 import test_video as tv
    Image = pytest.importorskip("PIL.Image")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_frames", os.path.join(os.path.dirname(__file__), "golden", "make_golden_frames.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for item in GOLD["big"]:
        seed, size, digest = str(item).split(":")
        a = mod.frame(int(seed), 1080, 1920)
        out = tv.preprocess_frames(torch.from_numpy(a).cuda(), (int(size), int(size)))
        assert hashlib.sha256(out[0].cpu().numpy().tobytes()).hexdigest() == digest
    # the reference's entry point: an image file (test_video.py:76-78) -> [1, 3, 256, 256]
    a = mod.frame(31, 240, 320)
    path = os.path.join(tmp_path, "frame.png")
    Image.fromarray(a).save(path)
    out = tv.preprocess_image(path)
    assert out.shape == (1, 3, 256, 256)
    assert np.array_equal(_bits(out[0]), mod.reference(a, (256, 256)).view(np.uint32))


def test_frames_argument_errors_and_result_dump():
    from ecologysemanticsegmentation_b200 import _native as nat
    from ecologysemanticsegmentation_b200 import test_video as tv
    with pytest.raises(nat.EcoLossError):
        tv.preprocess_frames(torch.zeros(4, 4, 3, dtype=torch.uint8), (2, 2))          # CPU tensor: no fallback
    with pytest.raises(TypeError):
        tv.preprocess_frames(torch.zeros(4, 4, 3, device="cuda"), (2, 2))
    with pytest.raises(ValueError):
        tv.preprocess_frames(torch.zeros(4, 4, 4, dtype=torch.uint8, device="cuda"), (2, 2))
    with pytest.raises(nat.EcoLossError):                                                 # one tile's patch beyond shared memory
        tv.preprocess_frames(torch.zeros(1, 4000, 4000, 3, dtype=torch.uint8, device="cuda"), (16, 16))
    # test_video.py:129-130: (output.squeeze().numpy() * 255).astype(np.uint8)
    p = torch.rand(1, 1, 64, 64, device="cuda")
    assert np.array_equal(tv.to_uint8_image(p).cpu().numpy(), (p.squeeze().cpu().numpy() * 255).astype(np.uint8))
