"""Golden vectors for the sequential model's test scoring, from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_seqtest.py

ess/test_multiclass_sequential_densenetloss.py:62,66,97-99 on seeded inputs (CPU, fp32): ``F.sigmoid`` ->
``return_union_sets_descending_order(out, reverse=True)`` (the reference's own function, utils/subsets_union.py:8-32, pulled
out of its file by ref_loader) -> per class ``dice_loss(out_c, lab_c, background_weight=0)`` (the reference's own
loss_functions.dice_loss).  Writes tests/golden/golden_seqtest.npz.
"""
import os
import sys

import numpy as np
import torch
from torch.nn import functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402


def main():
    assert ref_loader.available(), "needs /root/reference"
    lf, _, _ = ref_loader.load()
    union_cls, _, _ = ref_loader.load_adjacent()
    out = {}
    for tag, shape, seed in (("c3", (3, 3, 12, 12), 41), ("c4", (2, 4, 10, 14), 42), ("c2", (2, 2, 8, 8), 43)):
        torch.manual_seed(seed)
        z = torch.randn(shape) * 2
        lab = (torch.rand(shape) > 0.6).float()
        test_outputs = F.sigmoid(z)                                              # :62
        test_outputs = union_cls(test_outputs, reverse=True)                     # :66
        loss = [lf.dice_loss(test_outputs[:, idx:idx + 1, :, :], lab[:, idx:idx + 1, :, :], background_weight=0)
                for idx in range(lab.shape[1])]                                  # :97-98
        out[f"{tag}_z"], out[f"{tag}_lab"] = z.numpy(), lab.numpy()
        out[f"{tag}_dice"] = np.array([-float(v) for v in loss])                 # :99 accumulates x - l
        out[f"{tag}_out"] = test_outputs.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_seqtest.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
