"""Golden vectors for the primitives called with NON-default keyword parameters (focal gamma, Tversky alpha/beta,
focal-Dice gamma), produced by the UNMODIFIED reference (loss_functions.py:46,82,96) through ref_loader.

    python tests/golden/make_golden_shaped.py      (needs /root/reference; writes golden_shaped.npz)

Same two input tensors as golden_small.npz's ``prim_a`` / ``prim_b`` (seed 77); gradients w.r.t. both slots.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

CASES = {
    "focal_g2": lambda lf, a, b: lf.focal_loss(a, b, gamma=2.0),
    "focal_g07_bw": lambda lf, a, b: lf.focal_loss(a, b, gamma=0.7, factor=1, background_weight=0.4),
    "focal_g3_bw": lambda lf, a, b: lf.focal_loss(a, b, gamma=3, factor=0.5, background_weight=1),
    "twersky_a7b3": lambda lf, a, b: lf.twersky_loss(a, b, alpha=0.7, beta=0.3),
    "twersky_a2b8_bw": lambda lf, a, b: lf.twersky_loss(a, b, alpha=0.2, beta=0.8, background_weight=0.25),
    "focal_dice_g1": lambda lf, a, b: lf.focal_dice_coefficient(a, b, gamma=1.0),
    "focal_dice_g25_bw": lambda lf, a, b: lf.focal_dice_coefficient(a, b, gamma=2.5, background_weight=0.25),
}


def main():
    assert ref_loader.available(), "needs /root/reference"
    lf, _, _ = ref_loader.load()
    base = np.load(os.path.join(HERE, "golden_small.npz"))
    a0, b0 = torch.from_numpy(base["prim_a"]), torch.from_numpy(base["prim_b"])
    out = {}
    for name, fn in CASES.items():
        a = a0.clone().requires_grad_(True)
        b = b0.clone().requires_grad_(True)
        v = fn(lf, a, b)
        v.backward()
        out[f"{name}_val"] = np.array([float(v.detach())])
        out[f"{name}_ga"] = a.grad.numpy() if a.grad is not None else np.zeros_like(a0.numpy())
        out[f"{name}_gb"] = b.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_shaped.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
