"""Golden vectors of the frame pre-processing (ess/test_video.py:70-78) from Pillow + torchvision themselves
(run in the build container: Pillow 12.2.0, torchvision 0.26.0):

    python tests/golden/make_golden_frames.py

Writes tests/golden/golden_frames.npz: small cases WITH inputs and outputs, and SHA-256 digests of the outputs for seeded
1080p frames at the sizes of BASELINE.json's cfg5 (512) and of the reference's own transform (256)."""
import hashlib
import os

import numpy as np
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]   # test_video.py:74


def frame(seed, h, w):
    rs = np.random.RandomState(seed)
    a = (rs.rand(h, w, 3) * 255).astype(np.uint8)
    a[: max(1, h // 9), : max(1, w // 7)] = 255          # saturated and black regions: the clip8 edges
    a[h - max(1, h // 11):, :] = 0
    yy, xx = np.mgrid[0:h, 0:w]
    a[..., 1] = ((a[..., 1].astype(np.int32) + (xx * 255 // max(1, w - 1))) // 2).astype(np.uint8)   # a smooth ramp too
    return a


def reference(a, size):
    tf = transforms.Compose([transforms.Resize(size), transforms.ToTensor(), transforms.Normalize(mean=MEAN, std=STD)])
    return tf(Image.fromarray(a)).numpy()


def main():
    out = {}
    small = [(11, 90, 160, (32, 32)), (12, 37, 53, (64, 64)), (13, 100, 75, (40, 56)), (14, 64, 64, (64, 64)), (15, 19, 23, (7, 5))]
    for seed, h, w, size in small:
        a = frame(seed, h, w)
        out[f"in_{seed}"] = a
        out[f"out_{seed}"] = reference(a, size)
    big = []
    for seed, size in ((21, (512, 512)), (22, (256, 256))):
        a = frame(seed, 1080, 1920)
        r = reference(a, size)
        big.append((seed, size[0], hashlib.sha256(r.tobytes()).hexdigest()))
    out["big"] = np.array([f"{s}:{n}:{d}" for s, n, d in big])
    np.savez_compressed(os.path.join(HERE, "golden_frames.npz"), **out)
    print("wrote golden_frames.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
