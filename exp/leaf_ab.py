"""A/B timing of the cfg1 one-launch leaf step with a given library: python exp/leaf_ab.py [path/to/libecoloss.so]"""
import os, sys
sys.path.insert(0, ".")
from ecologysemanticsegmentation_b200 import _native
if len(sys.argv) > 1:
    _native.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from ecologysemanticsegmentation_b200 import fused
from ecologysemanticsegmentation_b200.synthetic import make_inputs
sets = [tuple(t.cuda() for t in make_inputs(54, 1, 256, 101 + k)) for k in range(4)]
outs = [torch.empty_like(z) for z, _ in sets]
step = fused.LeafLossStep(fused.loss_weights(bce=1.0, generalized_dice=1.0, twersky=1.0), doubling=1.0)
def run():
    for k in range(4):
        step(sets[k][0], sets[k][1], out=outs[k])
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    run()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    run()
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = []
for rep in range(3):
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 200 * 1e3)
print("cfg1 leaf step us:", [round(r, 2) for r in res])
