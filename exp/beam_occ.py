"""A/B of the beam kernel's CTAs/SM (exp/occN/libecoloss.so built with -DECO_BEAM_CTAS=N; default = the in-tree library)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ecologysemanticsegmentation_b200 import _native
if len(sys.argv) > 1:
    _native.LIB_PATH = os.path.join(ROOT, "exp", sys.argv[1], "libecoloss.so")
from ecologysemanticsegmentation_b200 import ops
from ecologysemanticsegmentation_b200.synthetic import make_inputs
z, g = make_inputs(54, 3, 1024, 103)
z, g = z.cuda(), g.cuda()
for nthr in (5, 19):
    thr = torch.linspace(0.8, 0.98, nthr, device="cuda")
    for _ in range(3): ops.dice_counts(z, g, thr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.dice_counts(z, g, thr)
    e1.record(); torch.cuda.synchronize()
    print(sys.argv[1] if len(sys.argv) > 1 else "in-tree", "n_thr", nthr, round(e0.elapsed_time(e1) / 20 * 1e3, 1), "us")
