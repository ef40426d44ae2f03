"""Beam (5..20 thresholds) kernel: counts vs single-threshold calls, NaN logits, timing at cfg3 (python exp/beam_check.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecologysemanticsegmentation_b200 import ops  # noqa: E402
from ecologysemanticsegmentation_b200.synthetic import make_inputs  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    n, c, s = 54, 3, 1024
    z, g = make_inputs(n, c, s, 103)
    zc, gc = z.cuda(), g.cuda()
    thr19 = torch.tensor(np.arange(0.8, 0.99, 0.01), dtype=torch.float32, device="cuda")
    counts19, soft19 = ops.dice_counts(zc, gc, thr19)
    ok = True
    for k in (0, 7, 18):
        c1, _ = ops.dice_counts(zc, gc, thr19[k:k + 1])
        same = torch.equal(c1[0], counts19[k])
        ok &= same
        print("threshold", k, "beam == single:", same)
    # NaN / inf logits: never above a threshold (NaN), always (inf)
    z2 = zc[:2].clone()
    z2[0, 0, 0, :7] = float("nan")
    z2[1, 1, 5, :9] = float("inf")
    z2[1, 2, 9, :11] = -float("inf")
    cb, _ = ops.dice_counts(z2, gc[:2], thr19)
    for k in (0, 18):
        c1, _ = ops.dice_counts(z2, gc[:2], thr19[k:k + 1])
        same = torch.equal(c1[0], cb[k])
        ok &= same
        print("nan/inf logits, threshold", k, "beam == single:", same)
    elems = n * c * s * s
    for nthr in (1, 2, 4, 5, 19):
        thr = torch.linspace(0.8, 0.98, nthr, device="cuda")
        us = timeit(lambda: ops.dice_counts(zc, gc, thr))
        print(f"n_thr={nthr}: {us:.1f} us  {8 * elems / us / 1e3:.0f} GB/s  frac {8 * elems / us / 1e3 / 6454:.3f}")
    gu8 = gc.to(torch.uint8)
    us = timeit(lambda: ops.dice_counts(zc, gu8, thr19))
    print(f"n_thr=19 u8 labels: {us:.1f} us  {5 * elems / us / 1e3:.0f} GB/s  frac {5 * elems / us / 1e3 / 6454:.3f}")
    cu8, _ = ops.dice_counts(zc, gu8, thr19)
    ok &= torch.equal(cu8, counts19)
    print("u8 == f32 labels:", torch.equal(cu8, counts19))
    print("ALL OK" if ok else "MISMATCH")


if __name__ == "__main__":
    main()
