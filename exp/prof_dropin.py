"""Host-side profile of the drop-in autograd path (development helper)."""
import cProfile, pstats, sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import ecologysemanticsegmentation_b200 as eco
from ecologysemanticsegmentation_b200.synthetic import make_inputs
z, g = make_inputs(54, 3, 256, 102)
z, g = z.cuda().requires_grad_(True), g.cuda()
def step():
    z.grad = None
    ce, bce, fl, dice, gdice, tw, fd = eco.losses_fn(z, g, True, from_logits=True)
    (1.0 * fd + 1.0 * bce + 1.0 * (gdice + tw)).backward()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host us/step", (t1 - t0) / 50 * 1e6, "incl. drain", (t2 - t0) / 50 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
