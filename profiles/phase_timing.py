"""Per-kernel timing of the composite path at cfg2 (CUDA events, rotating buffer sets):
    python profiles/phase_timing.py [--iters 200]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecologysemanticsegmentation_b200 import fused, ops  # noqa: E402
from ecologysemanticsegmentation_b200.loss_composite import DEFAULT_RATIOS, composite3_leaf_scales, draw_pair_weights  # noqa: E402
from ecologysemanticsegmentation_b200.synthetic import make_config  # noqa: E402


def timeit(fn, iters, nsets):
    for i in range(10):
        fn(i % nsets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nsets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--cfg", default="cfg2")
    args = ap.parse_args()
    nsets = 4
    z0, g0 = make_config(args.cfg)
    sets = [(z0.cuda() + 0.01 * k, g0.cuda().clone()) for k in range(nsets)]
    outs = [torch.empty_like(z) for z, _ in sets]
    np.random.seed(0)
    scales = composite3_leaf_scales(draw_pair_weights(DEFAULT_RATIOS, False))
    up = torch.tensor(fused.loss_weights(bce=1, generalized_dice=1, twersky=1, focal_dice=1), dtype=torch.float32, device="cuda")
    step = fused.CompositeLossStep(fused.loss_weights(bce=1, generalized_dice=1, twersky=1, focal_dice=1))
    acc = ops.composite3_stats(*sets[0], True)
    losses, jac, _ = ops.composite3_finalize(acc, scales)
    res = {}
    res["stats"] = timeit(lambda i: ops.composite3_stats(*sets[i], True), args.iters, nsets)
    res["finalize"] = timeit(lambda i: ops.composite3_finalize(acc, scales), args.iters, nsets)
    res["grad"] = timeit(lambda i: ops.composite3_grad(*sets[i], True, jac, up, out=outs[i]), args.iters, nsets)
    res["fused"] = timeit(lambda i: step(*sets[i], out=outs[i]), args.iters, nsets)
    up0 = torch.tensor(fused.loss_weights(generalized_dice=1, twersky=1, focal_dice=1), dtype=torch.float32, device="cuda")
    res["grad_no_bce"] = timeit(lambda i: ops.composite3_grad(*sets[i], True, jac, up0, out=outs[i]), args.iters, nsets)
    n, c, h, w = z0.shape
    print({k: round(v, 2) for k, v in res.items()}, "us;  shape", tuple(z0.shape),
          " fused GB/s (12 B/elem):", round(12 * n * c * h * w / res["fused"] / 1e3, 1))


if __name__ == "__main__":
    main()
