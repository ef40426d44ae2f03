"""Adjacent kernels that the bench does not time: the in-place label union / un-union and the soft-label CE."""
import sys
import torch
sys.path.insert(0, ".")
import ecologysemanticsegmentation_b200 as eco
from ecologysemanticsegmentation_b200 import subsets_union, train_multiclass
def timed(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for shape in ((54, 3, 512, 512), (54, 3, 1024, 1024)):
    g = (torch.rand(shape, device="cuda") > 0.6).float()
    p = torch.rand(shape, device="cuda")
    n = g.numel()
    t = timed(lambda: subsets_union.return_union_sets_descending_order(g))
    print(f"{shape} label union (class dim, in place): {t:7.1f} us  {8*n*2/3/t/1e3:6.0f} GB/s over the 2 touched planes (r+w), {8*n/t/1e3:6.0f} GB/s if all 3 counted")
    t = timed(lambda: subsets_union.return_union_sets_descending_order(p, reverse=True))
    print(f"{shape} prediction un-union (in place):     {t:7.1f} us")
    t = timed(lambda: train_multiclass.return_union_sets_descending_order(g))
    print(f"{shape} label union (batch dim twin):       {t:7.1f} us")
    pr = p.clone().requires_grad_(True)
    def ce():
        pr.grad = None
        eco.loss_functions.cross_entropy_loss(g, pr).backward()
    t = timed(ce, 10)
    print(f"{shape} soft-label CE fwd+bwd:              {t:7.1f} us  ({(8+12)*n/t/1e3:6.0f} GB/s at 8 + 12 B/element)")
