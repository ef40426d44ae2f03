"""Host time of one fused-step call (development helper)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from ecologysemanticsegmentation_b200 import fused
from ecologysemanticsegmentation_b200.synthetic import make_inputs
z, g = make_inputs(54, 3, 256, 102)
z, g = z.cuda(), g.cuda()
out = torch.empty_like(z)
step = fused.CompositeLossStep([0, 1, 0, 0, 1, 1, 1])
for _ in range(5): step(z, g, out=out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step(z, g, out=out)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("fused step: host us/call", (t1 - t0) / 200 * 1e6, " wall us/step", (t2 - t0) / 200 * 1e6)
