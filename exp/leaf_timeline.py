"""In-kernel timeline of pair_fused_kernel (cfg1 shape and the cfg2 shape with 3 leaves).  Needs exp/tl/libecoloss_tl.so
(eco_leaf.cu compiled with -DECO_LEAF_TIMELINE; see the build lines in DESIGN.md / exp/README).  Prints, per phase, the
mean and max over CTAs of the time since the first CTA's start, in us."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ecologysemanticsegmentation_b200 import _native  # noqa: E402

_native.LIB_PATH = os.path.join(ROOT, "exp", "tl", "libecoloss_tl.so")
from ecologysemanticsegmentation_b200 import fused  # noqa: E402
from ecologysemanticsegmentation_b200.synthetic import make_inputs  # noqa: E402

NAMES = ["start", "pass-1 loop done", "arrived (+ last-CTA sum)", "coefficients in", "pass 2 done", "last: sums ready", "last: handed over", "last: closed forms done"]


def main():
    L = _native.lib()
    L.eco_debug_leaf_timeline.restype = C.c_int
    L.eco_debug_leaf_timeline.argtypes = [C.c_void_p, C.c_int]
    for (n, c, s) in ((54, 1, 256), (54, 3, 256)):
        sets = [tuple(t.cuda() for t in make_inputs(n, c, s, 101 + k)) for k in range(6)]
        step = fused.LeafLossStep(fused.loss_weights(bce=1.0, generalized_dice=1.0, twersky=1.0), doubling=1.0)
        for k in range(12):
            step(*sets[k % 6])
        torch.cuda.synchronize()
        buf = np.zeros(4096 * 8, dtype=np.uint64)
        assert L.eco_debug_leaf_timeline(buf.ctypes.data, buf.size) == 0
        tl = buf.reshape(4096, 8).astype(np.int64)
        tl = tl[tl[:, 0] > 0]
        t0 = tl[:, 0].min()
        print(f"shape {n}x{c}x{s}x{s}: {len(tl)} CTAs")
        for k in (0, 1, 2, 3, 4):
            v = (tl[:, k] - t0) / 1e3
            print(f"  {NAMES[k]:28s} mean {v.mean():6.2f}  max {v.max():6.2f} us")
        last = tl[tl[:, 5] > 0]
        for k in (5, 7, 6):
            v = (last[:, k] - t0) / 1e3
            print(f"  {NAMES[k]:28s} mean {v.mean():6.2f}  max {v.max():6.2f} us")


if __name__ == "__main__":
    main()
