"""Seeded synthetic inputs for the five BASELINE.json configs.

Tensors are generated on the CPU with ``torch.Generator().manual_seed(seed)`` so that the CPU
oracle and the GPU kernels see identical bits.  Logits ~ N(0,1) fp32; masks are fp32 {0,1},
nested from one uniform field so that whole_body >= (ventral+dorsal) >= dorsal, matching the
reference's ``relative_set_ratios`` (ecology_semantic_segmentation/loss_composite.py:21).
"""
from __future__ import annotations

import torch

RELATIVE_SET_RATIOS = (1.0, 0.43197708, 0.22319692)

# name -> (seed, N, C, S)
CONFIGS = {
    "cfg1": (101, 54, 1, 256),    # whole_body single-class leaf
    "cfg2": (102, 54, 3, 256),    # 3-organ composite loss  (the headline workload)
    "cfg3": (103, 54, 3, 1024),   # thresholded Dice scoring
    "cfg4": (104, 432, 3, 512),   # composite loss, 54 images per GPU x 8
    "cfg5": (105, 64, 3, 512),    # one batch of the frame stream (per GPU)
}


def make_inputs(n, c, s, seed, logit_scale=1.0, nested=True, pin=False):
    """Returns (logits f32[n,c,s,s], masks f32[n,c,s,s]) on the CPU."""
    gen = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, c, s, s, generator=gen, dtype=torch.float32)
    if logit_scale != 1.0:
        logits = logits * logit_scale
    if nested:
        u = torch.rand(n, 1, s, s, generator=gen, dtype=torch.float32)
        ratios = [RELATIVE_SET_RATIOS[k] if k < len(RELATIVE_SET_RATIOS) else RELATIVE_SET_RATIOS[-1] / (k + 1)
                  for k in range(c)]
        masks = torch.cat([(u < 0.5 * r) for r in ratios], dim=1).to(torch.float32)
    else:
        masks = (torch.rand(n, c, s, s, generator=gen, dtype=torch.float32) > 0.5).to(torch.float32)
    if pin:
        logits, masks = logits.pin_memory(), masks.pin_memory()
    return logits, masks


def make_config(name, n=None, **kw):
    seed, n0, c, s = CONFIGS[name]
    return make_inputs(n0 if n is None else n, c, s, seed, **kw)
