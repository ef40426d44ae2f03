"""Generate golden vectors from the UNMODIFIED reference (run in the build container only):

    python tests/golden/make_golden.py

Imports /root/reference's own loss_functions.py / loss_composite.py / train_multiclass.losses_fn through
ref_loader.py, evaluates them on seeded inputs (CPU, fp32) and writes
    tests/golden/golden_small.npz   small cases WITH inputs (usable anywhere)
    tests/golden/golden_meta.json   scalar outputs, gradient digests of the full-size configs, input hashes
The reference cannot travel to the GPU box; these files can.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_loader  # noqa: E402
from ecologysemanticsegmentation_b200.synthetic import make_config, make_inputs  # noqa: E402

UP_ALL = [0.3, 1.0, 0.7, 0.2, 1.0, 0.5, 1.0]
UP_CFG2 = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0]   # bce + gdice + twersky + focal_dice
UP_CFG1 = [0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 0.0]   # bce + gdice + twersky (epoch<1000 weights)
SAMPLE_IDX_SEED = 999


def sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def grad_digest(g):
    g = g.double()
    flat = g.flatten()
    rng = np.random.RandomState(SAMPLE_IDX_SEED)
    idx = rng.randint(0, flat.numel(), size=256)
    return {"l2": float(flat.norm()), "max_abs": float(flat.abs().max()), "sum": float(flat.sum()),
            "sample_idx": idx.tolist(), "sample_val": flat[idx].tolist()}


def run_losses(fn, x, g, up, *args, **kw):
    x = x.clone().requires_grad_(True)
    losses = fn(x, g, *args, **kw)
    total = sum(float(w) * l for w, l in zip(up, losses) if w != 0.0)
    total.backward()
    return [float(v.detach()) for v in losses], x.grad.detach().clone()


def main():
    assert ref_loader.available(), "needs /root/reference"
    lf, lc, tm = ref_loader.load()
    small, meta = {}, {"torch": torch.__version__, "numpy": np.__version__, "cases": {}}

    # ---- small cases with inputs -------------------------------------------------------------------
    torch.manual_seed(20240)
    z = torch.randn(2, 3, 8, 8)
    p = torch.sigmoid(z)
    u = torch.rand(2, 1, 8, 8)
    g_nested = torch.cat([(u < 0.5), (u < 0.3), (u < 0.15)], 1).float()
    g_iid = (torch.rand(2, 3, 8, 8) > 0.5).float()
    small["p"], small["z"], small["g_nested"], small["g_iid"] = p.numpy(), z.numpy(), g_nested.numpy(), g_iid.numpy()

    for bw in (0, 0.5):
        l, gr = run_losses(lc.losses_fn, p[:, :1], g_iid[:, :1], UP_ALL, False, bw)
        small[f"leaf_lc_bw{bw}_losses"], small[f"leaf_lc_bw{bw}_grad"] = np.array(l), gr.numpy()
        l, gr = run_losses(tm, p[:, :1], g_iid[:, :1], UP_ALL, False, bw)
        small[f"leaf_tm_bw{bw}_losses"], small[f"leaf_tm_bw{bw}_grad"] = np.array(l), gr.numpy()
    l, gr = run_losses(lc.losses_fn, p, g_iid, UP_ALL, False, 0.7)  # bw must be ignored for C>1
    small["plain_lc_losses"], small["plain_lc_grad"] = np.array(l), gr.numpy()
    l, gr = run_losses(tm, p, g_iid, UP_ALL, True, 0.7)  # composite flag ignored for C>1 in train_multiclass
    small["plain_tm_losses"], small["plain_tm_grad"] = np.array(l), gr.numpy()
    for name, g in (("nested", g_nested), ("iid", g_iid)):
        np.random.seed(0)
        l, gr = run_losses(lc.losses_fn, p, g, UP_ALL, True)
        small[f"comp_{name}_losses"], small[f"comp_{name}_grad"] = np.array(l), gr.numpy()
    np.random.seed(123)
    l, gr = run_losses(lc.losses_fn, p, g_nested, UP_CFG2, True, 0, True)
    small["comp_es_losses"], small["comp_es_grad"] = np.array(l), gr.numpy()
    small["comp_es_rng_after"] = np.random.get_state()[1].copy()

    # primitives on two continuous tensors, gradients w.r.t. both slots
    torch.manual_seed(77)
    a0 = torch.rand(2, 3, 8, 8) * 0.96 + 0.02
    b0 = torch.rand(2, 3, 8, 8) * 0.96 + 0.02
    small["prim_a"], small["prim_b"] = a0.numpy(), b0.numpy()

    def prim(name, fn):
        a = a0.clone().requires_grad_(True)
        b = b0.clone().requires_grad_(True)
        out = fn(a, b)
        outs = out if isinstance(out, tuple) else (out,)
        sum((k + 1.0) * o for k, o in enumerate(outs)).backward()
        small[f"prim_{name}_val"] = np.array([float(o.detach()) for o in outs])
        small[f"prim_{name}_ga"] = a.grad.numpy() if a.grad is not None else np.zeros_like(a0.numpy())
        small[f"prim_{name}_gb"] = b.grad.numpy() if b.grad is not None else np.zeros_like(b0.numpy())

    prim("bce", lambda a, b: lf.cross_entropy_loss(a, b, bce=True))
    prim("softce", lambda a, b: lf.cross_entropy_loss(a, b))
    prim("softce_bw", lambda a, b: lf.cross_entropy_loss(a, b, background_weight=0.3))
    prim("focal", lambda a, b: lf.focal_loss(a, b))
    prim("focal_bw", lambda a, b: lf.focal_loss(a, b, factor=1, background_weight=0.4))
    prim("dice", lambda a, b: lf.dice_loss(a, b))
    prim("dice_bw0", lambda a, b: lf.dice_loss(a, b, background_weight=0))
    prim("gdice", lambda a, b: lf.dice_loss(a, b, generalized=True, background_weight=0.5))
    prim("twersky", lambda a, b: lf.twersky_loss(a, b, background_weight=0.25))
    prim("focal_dice", lambda a, b: lf.focal_dice_coefficient(a, b, background_weight=0.25))
    prim("cls_dice", lambda a, b: lf.classification_dice_loss(a, b))
    prim("cls_dice_f10", lambda a, b: lf.classification_dice_loss(a, b, factor=10, background_weight=0))

    # evaluation: reference thresholding + dice_loss per class, and exact counts of the thresholded tensors
    torch.manual_seed(5)
    ze = torch.randn(2, 3, 16, 16) * 2
    le = (torch.rand(2, 3, 16, 16) > 0.6).float()
    small["eval_z"], small["eval_lab"] = ze.numpy(), le.numpy()
    for thr in (None, 0.8, 0.9):
        out = torch.sigmoid(ze)
        if thr is not None:
            out[out > thr] = 1
            out[out != 1] = 0
        dice = [-float(lf.dice_loss(out[:, c:c + 1], le[:, c:c + 1], background_weight=0)) for c in range(3)]
        small[f"eval_dice_{thr}"] = np.array(dice)
        if thr is not None:
            small[f"eval_counts_{thr}"] = np.array(
                [[int((out[:, c].long() * le[:, c].long()).sum()), int(out[:, c].long().sum()), int(le[:, c].long().sum())]
                 for c in range(3)], dtype=np.int64)

    # ---- adjacent steps: label union / un-union, sequential-variant losses_fn ---------------------------------
    union_cls, union_bat, seq = ref_loader.load_adjacent()
    torch.manual_seed(31)
    ann = (torch.rand(5, 4, 6, 6) > 0.6).float()
    prob = torch.rand(5, 4, 6, 6)
    small["union_ann"], small["union_prob"] = ann.numpy(), prob.numpy()
    for tag, ex in (("e0", [0]), ("e02", [0, 2]), ("none", [])):
        small[f"union_cls_fwd_{tag}"] = union_cls(ann.clone(), ex).numpy()
        small[f"union_cls_rev_{tag}"] = union_cls(prob.clone(), ex, reverse=True).numpy()
        small[f"union_bat_fwd_{tag}"] = union_bat(ann.clone(), ex).numpy()
    l, gr = run_losses(seq, p, g_nested, UP_ALL)
    small["seq_losses"], small["seq_grad"] = np.array(l), gr.numpy()
    l, gr = run_losses(seq, p[:, :1], g_iid[:, :1], UP_ALL, False, 0.5)
    small["seq_c1_losses"], small["seq_c1_grad"] = np.array(l), gr.numpy()

    # ---- SURVEY.md 8(c) known-answer values (seed 0, 4x3x64x64) -----------------------------------------
    torch.manual_seed(0)
    pk = torch.sigmoid(torch.randn(4, 3, 64, 64))
    gk = (torch.rand(4, 3, 64, 64) > 0.5).float()
    kat = {"p_sha": sha(pk), "g_sha": sha(gk)}
    kat["lc_c1"] = [float(v) for v in lc.losses_fn(pk[:, :1], gk[:, :1])]
    kat["lc_c3"] = [float(v) for v in lc.losses_fn(pk, gk)]
    np.random.seed(0)
    kat["lc_c3_composite"] = [float(v) for v in lc.losses_fn(pk, gk, composite_set_theory=True)]
    kat["tm_c3"] = [float(v) for v in tm(pk, gk)]
    meta["cases"]["kat_seed0_4x3x64x64"] = kat

    # ---- full-size configs: outputs + gradient digests ---------------------------------------------------
    z1, g1 = make_config("cfg1")
    p1 = torch.sigmoid(z1)
    for bw in (0, 0.5):
        l, gr = run_losses(lc.losses_fn, p1, g1, UP_CFG1, False, bw)
        meta["cases"][f"cfg1_lc_bw{bw}"] = {"z_sha": sha(z1), "g_sha": sha(g1), "p_sha": sha(p1), "upstream": UP_CFG1,
                                            "losses": l, "grad_wrt_p": grad_digest(gr)}
        l, gr = run_losses(tm, p1, g1, UP_CFG1, False, bw)
        meta["cases"][f"cfg1_tm_bw{bw}"] = {"upstream": UP_CFG1, "losses": l, "grad_wrt_p": grad_digest(gr)}
    z2, g2 = make_config("cfg2")
    p2 = torch.sigmoid(z2)
    np.random.seed(0)
    l, gr = run_losses(lc.losses_fn, p2, g2, UP_CFG2, True)
    meta["cases"]["cfg2_composite"] = {"z_sha": sha(z2), "g_sha": sha(g2), "p_sha": sha(p2), "upstream": UP_CFG2,
                                       "np_seed": 0, "losses": l, "grad_wrt_p": grad_digest(gr)}
    l, gr = run_losses(lc.losses_fn, p2, g2, UP_CFG2, False)
    meta["cases"]["cfg2_plain"] = {"upstream": UP_CFG2, "losses": l, "grad_wrt_p": grad_digest(gr)}
    # gradient w.r.t. logits through the CPU sigmoid, for orientation (same-device parity is the gate)
    zz = z2.clone().requires_grad_(True)
    np.random.seed(0)
    ls = lc.losses_fn(torch.sigmoid(zz), g2, True)
    sum(float(w) * v for w, v in zip(UP_CFG2, ls) if w).backward()
    meta["cases"]["cfg2_composite"]["grad_wrt_z_cpu_sigmoid"] = grad_digest(zz.grad)

    # cfg3-shaped scoring on a reduced batch (8 images of 1024^2 keeps the CPU run short)
    z3, g3 = make_inputs(8, 3, 1024, 103)
    ev = {"z_sha": sha(z3), "g_sha": sha(g3)}
    for thr in (None, 0.8, 0.9):
        out = torch.sigmoid(z3)
        if thr is not None:
            out[out > thr] = 1
            out[out != 1] = 0
            ev[f"counts_{thr}"] = [[int((out[:, c].long() * g3[:, c].long()).sum()), int(out[:, c].long().sum()),
                                    int(g3[:, c].long().sum())] for c in range(3)]
        ev[f"dice_{thr}"] = [-float(lf.dice_loss(out[:, c:c + 1], g3[:, c:c + 1], background_weight=0)) for c in range(3)]
    meta["cases"]["cfg3_n8"] = ev

    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **small)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", len(small), "arrays and", len(meta["cases"]), "meta cases")


if __name__ == "__main__":
    main()
