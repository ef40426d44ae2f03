"""Achieved HBM GB/s of the memory-bound kernels (CUDA events, rotating buffer sets so nothing is L2-resident):
    python profiles/kernel_rooflines.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecologysemanticsegmentation_b200 import ops  # noqa: E402
from ecologysemanticsegmentation_b200.synthetic import make_inputs  # noqa: E402

PEAK = 6454.0


def timeit(fn, iters, nsets):
    for i in range(5):
        fn(i % nsets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % nsets)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def report(name, secs, nbytes):
    gbs = nbytes / secs / 1e9
    print(json.dumps({"kernel": name, "us": round(secs * 1e6, 2), "algorithmic_MB": round(nbytes / 1e6, 1),
                      "GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3)}))


def main():
    up = torch.tensor([0, 1, 0, 0, 1, 1, 1], dtype=torch.float32, device="cuda")
    # --- leaf engine, cfg2 shape (3 channel leaves) and cfg1 shape (one leaf), probabilities in ---
    for name, (n, c, s), nsets in (("cfg2-shape", (54, 3, 256), 6), ("cfg1-shape", (54, 1, 256), 12), ("cfg4-shard", (54, 3, 512), 3)):
        z, g = make_inputs(n, c, s, 7)
        sets = [(torch.sigmoid(z).cuda() * (1 - 0.001 * k), g.cuda().clone()) for k in range(nsets)]
        outs = [torch.empty_like(a) for a, _ in sets]
        elems = n * c * s * s
        t = timeit(lambda i: ops.pair_stats(sets[i][1], sets[i][0], 0), 100, nsets)
        report(f"pair_stats {name}", t, 8 * elems)
        sums = ops.pair_stats(sets[0][1], sets[0][0], 0)
        _, _, jac = ops.pair_finalize(sums, 0.0, [2.0] * c)
        t = timeit(lambda i: ops.pair_grad(sets[i][1], sets[i][0], 0, jac, up, False, True), 100, nsets)
        report(f"pair_grad(d/db) {name}", t, 12 * elems)
        t = timeit(lambda i: ops.pair_stats(sets[i][1], sets[i][0], ops.nat.FLAG_B_LOGIT), 100, nsets)
        report(f"pair_stats from logits {name}", t, 8 * elems)
    # --- scoring, cfg3 shape ---
    n, c, s = 54, 3, 1024
    z, g = make_inputs(n, c, s, 103)
    zc, gc = z.cuda(), g.cuda()
    elems = n * c * s * s
    for nthr in (0, 1, 4, 19):
        thr = None if nthr == 0 else torch.linspace(0.8, 0.98, nthr, device="cuda")
        t = timeit(lambda i: ops.dice_counts(zc, gc, thr), 20, 1)
        report(f"dice_counts cfg3 n_thr={nthr}", t, 8 * elems)
    # --- byte masks for the result dumps (sigmoid -> threshold -> *255 -> uint8), cfg3 shape: 4 B in + 1 B out ---
    for label, thr in (("sigmoid", None), ("threshold 0.8", 0.8)):
        t = timeit(lambda i: ops.masks_u8(zc, thr), 20, 1)
        report(f"masks_u8 cfg3 {label}", t, 5 * elems)
    t = timeit(lambda i: ops.masks_u8(gc, None, True), 20, 1)
    report("masks_u8 cfg3 labels", t, 5 * elems)


if __name__ == "__main__":
    main()
