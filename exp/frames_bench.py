"""Frame pre-processing kernel: batch of 1080p frames -> 512 / 256 (python exp/frames_bench.py)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecologysemanticsegmentation_b200 import test_video as tv  # noqa: E402


def main():
    n = 64
    rs = np.random.RandomState(0)
    frames = torch.from_numpy((rs.rand(n, 1080, 1920, 3) * 255).astype(np.uint8)).cuda()
    for size in ((512, 512), (256, 256)):
        out = torch.empty((n, 3) + size, dtype=torch.float32, device="cuda")
        for _ in range(3):
            tv.preprocess_frames(frames, size, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            tv.preprocess_frames(frames, size, out=out)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        nbytes = frames.numel() + out.numel() * 4
        print(f"{n} x 1080p -> {size}: {us:.1f} us per batch, {nbytes / us / 1e3:.0f} GB/s algorithmic ({nbytes / us / 1e3 / 6454:.3f} of peak)")
    # the reference's way for one frame on this box's CPU: Pillow resize + ToTensor + Normalize
    try:
        from PIL import Image
        from torchvision import transforms
        tf = transforms.Compose([transforms.Resize((512, 512)), transforms.ToTensor(),
                                 transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        im = Image.fromarray(frames[0].cpu().numpy())
        t0 = time.perf_counter()
        for _ in range(5):
            tf(im)
        print(f"Pillow + torchvision on the host: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per frame")
    except Exception as exc:  # noqa: BLE001
        print("no Pillow / torchvision here:", exc)


if __name__ == "__main__":
    main()
