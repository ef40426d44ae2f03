"""A few launches of the beam kernel for ncu (f32 labels, then u8 labels)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecologysemanticsegmentation_b200 import ops  # noqa: E402
from ecologysemanticsegmentation_b200.synthetic import make_inputs  # noqa: E402

z, g = make_inputs(54, 3, 1024, 103)
zc, gc = z.cuda(), g.cuda()
gu8 = gc.to(torch.uint8)
thr19 = torch.tensor(np.arange(0.8, 0.99, 0.01), dtype=torch.float32, device="cuda")
for _ in range(3):
    ops.dice_counts(zc, gc, thr19)
for _ in range(3):
    ops.dice_counts(zc, gu8, thr19)
torch.cuda.synchronize()
