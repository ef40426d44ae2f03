"""Quick check + timing of the fused composite step on the GPU box (development helper)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from ecologysemanticsegmentation_b200 import fused
from ecologysemanticsegmentation_b200.synthetic import make_inputs
from oracle import torch_port as tp

UP = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]
torch.cuda.set_device(0)
for shape in [(2, 16), (5, 64), (54, 256)]:
    z, g = make_inputs(shape[0], 3, shape[1], 7)
    z, g = z.cuda(), g.cuda()
    np.random.seed(0)
    step = fused.CompositeLossStep(UP)
    l, d = step(z, g)
    torch.cuda.synchronize()
    zr = z.clone().requires_grad_(True)
    np.random.seed(0)
    ref = tp.losses_composite(torch.sigmoid(zr), g, True)
    sum(w * v for w, v in zip(UP, ref) if w).backward()
    rl = torch.stack([v.detach() for v in ref])
    el = float(((l[1:] - rl[1:]).abs() / rl[1:].abs()).max())
    eg = float((d - zr.grad).abs().max() / zr.grad.abs().max())
    l8, d8 = step(z, g.to(torch.uint8))
    print(shape, "loss rel", el, "grad max", eg, "u8 identical", bool(torch.equal(l8, l) and torch.equal(d8, d)), flush=True)

n, s = 54, 256
sets = [tuple(t.cuda() for t in make_inputs(n, 3, s, 100 + k)) for k in range(4)]
sets8 = [(z, g.to(torch.uint8)) for z, g in sets]
outs = [torch.empty_like(z) for z, _ in sets]
step = fused.CompositeLossStep(UP)
for name, ss in (("f32 labels", sets), ("u8 labels", sets8)):
    for i in range(10):
        step(ss[i % 4][0], ss[i % 4][1], out=outs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200):
        step(ss[i % 4][0], ss[i % 4][1], out=outs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    print(name, "us/step", e0.elapsed_time(e1) / 200 * 1e3, flush=True)
