"""Scoring kernels with byte masks (5 B/element) against float32 masks (8 B/element) at cfg3."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ecologysemanticsegmentation_b200 import ops
from ecologysemanticsegmentation_b200.synthetic import make_inputs
z, g = make_inputs(54, 3, 1024, 103)
z, g = z.cuda(), g.cuda()
g8 = g.to(torch.uint8)
def timed(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
n = z.numel()
for name, thr in (("soft", None), ("1 thr", torch.tensor([0.8], device="cuda")), ("4 thr", torch.tensor([0.5, 0.6, 0.7, 0.8], device="cuda")),
                  ("19 thr", torch.tensor(np.arange(0.8, 0.99, 0.01), dtype=torch.float32, device="cuda"))):
    a = timed(lambda: ops.dice_counts(z, g, thr)); b = timed(lambda: ops.dice_counts(z, g8, thr))
    print(f"{name:7s} f32 labels {a:7.1f} us = {8*n/a/1e3:6.0f} GB/s   u8 labels {b:7.1f} us = {5*n/b/1e3:6.0f} GB/s", flush=True)
