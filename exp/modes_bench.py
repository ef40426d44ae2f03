"""Step times of the one-launch composite kernel's modes at cfg2 (python exp/modes_bench.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecologysemanticsegmentation_b200 import ops  # noqa: E402
from ecologysemanticsegmentation_b200.fused import CompositeLossStep  # noqa: E402
from ecologysemanticsegmentation_b200.synthetic import make_inputs  # noqa: E402

UP = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]


def timeit(fn, iters=200, nsets=4):
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
    n, c, s = 54, 3, 256
    z, g = make_inputs(n, c, s, 102, nested=True)
    zs = [(z * (1 + 0.01 * k)).cuda() for k in range(4)]
    ps = [torch.sigmoid(v) for v in zs]
    gs = [g.cuda().clone() for _ in range(4)]
    outs = [torch.empty_like(v) for v in zs]
    np.random.seed(0)
    for name, xs, fl in (("logits", zs, True), ("probabilities", ps, False)):
        step = CompositeLossStep(UP, from_logits=fl)
        ents = [ops.PreparedComposite3(xs[k], gs[k], step.scales, step.upstream, fl) for k in range(4)]
        us = timeit(lambda i: ents[i].run(out=outs[i]))
        print(f"{name}: full step {us:.1f} us")
        us = timeit(lambda i: ents[i].run(no_grad=True))
        print(f"{name}: no-grad step {us:.1f} us")
    us = timeit(lambda i: ops.composite3_finalize(ops.composite3_stats(zs[i], gs[i], True), [1.0] * 21))
    print(f"first-generation statistics + finalize (what no_grad ran before): {us:.1f} us")


if __name__ == "__main__":
    main()
