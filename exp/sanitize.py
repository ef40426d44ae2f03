"""One small call of every kernel family, for compute-sanitizer (memcheck / initcheck / racecheck) on the GPU box:
    compute-sanitizer --tool memcheck --error-exitcode 7 python exp/sanitize.py
Shapes are small and deliberately ragged (tails, unaligned planes, strided channel slices)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import ecologysemanticsegmentation_b200 as eco  # noqa: E402
from ecologysemanticsegmentation_b200 import (fused, ops, subsets_union, test_multiclass, test_multiclass_sequential_densenetloss,  # noqa: E402
                                              test_video, train_multiclass, train_multiclass_sequential_densenetloss as seq_loss)
from ecologysemanticsegmentation_b200.synthetic import make_inputs  # noqa: E402

torch.cuda.set_device(0)
UP = fused.loss_weights(bce=1.0, generalized_dice=1.0, twersky=1.0, focal_dice=1.0)
UP_ALL = [0.3, 1.0, 0.7, 0.2, 1.0, 0.5, 1.0]


def combine(losses, up):
    return sum(w * l for w, l in zip(up, losses) if w)


def main():
    for n, s in ((2, 16), (3, 64), (5, 36)):
        z, g = make_inputs(n, 3, s, 11 + s)
        z, g = z.cuda(), g.cuda()
        # fused composite step: fp32 / byte labels, union at load, probabilities, no-grad, bf16
        np.random.seed(0)
        step = fused.CompositeLossStep(UP)
        step(z, g)
        step(z, g.to(torch.uint8))
        fused.CompositeLossStep(UP, union_labels=True)(z, g)
        step(z.bfloat16(), g)
        # drop-in autograd paths (composite fast path twice: miss, hit; plain 3-organ; one organ; generic organ count)
        for _ in range(2):
            zz = z.clone().requires_grad_(True)
            np.random.seed(0)
            combine(eco.losses_fn(torch.sigmoid(zz), g, True), UP).backward()
            zz = z.clone().requires_grad_(True)
            combine(train_multiclass.losses_fn(torch.sigmoid(zz), g, False, 0, False), UP).backward()
            zz = z.clone().requires_grad_(True)
            combine(train_multiclass.losses_fn(torch.sigmoid(zz[:, :1]), g[:, :1], False, 0.5, False), UP_ALL).backward()
            zz = z.clone().requires_grad_(True)
            combine(seq_loss.losses_fn(torch.sigmoid(zz), g), UP_ALL).backward()
        with torch.no_grad():
            np.random.seed(0)
            eco.losses_fn(torch.sigmoid(z), g, True)
        z4 = torch.randn(n, 4, s, s).cuda().requires_grad_(True)
        g4 = (torch.rand(n, 4, s, s) > 0.5).float().cuda()
        np.random.seed(0)
        combine(eco.losses_fn(torch.sigmoid(z4), g4, True, relative_set_ratios=[1.0, 0.6, 0.4, 0.2]), UP).backward()
        # one-launch steps of the extension API
        fused.MulticlassLossStep(UP)(z, g)
        fused.LeafLossStep(UP_ALL, doubling=1.0, background_weight=0.5)(z[:, :1].contiguous(), g[:, :1].contiguous())
        fused.LeafLossStep(UP_ALL)(z, g)
        # stand-alone primitives (soft-label CE included)
        p = torch.sigmoid(z).requires_grad_(True)
        (eco.loss_functions.dice_loss(g, p) + eco.loss_functions.focal_loss(g, p) + eco.loss_functions.cross_entropy_loss(g, p)
         + eco.loss_functions.twersky_loss(g, p, alpha=0.7) + eco.loss_functions.cross_entropy_loss(g, p, bce=True)).backward()
        # scoring: soft, 1 / 4 / 19 thresholds, byte masks, un-union, stream scorer, result dumps
        for thr in (None, [0.8], [0.5, 0.6, 0.7, 0.8], list(np.arange(0.8, 0.99, 0.01))):
            test_multiclass.score_batch(z, g, thr if thr is None or len(thr) > 1 else thr[0])
            test_multiclass.score_batch(z, g.to(torch.uint8), thr if thr is None or len(thr) > 1 else thr[0])
        test_multiclass_sequential_densenetloss.score_batch(z, g)
        test_multiclass_sequential_densenetloss.score_batch(z4.detach(), g4)
        sc = test_multiclass.StreamScorer(3, 2, 0.8)
        sc.add(z, g); sc.add(z * 0.5, g); sc.result()
        test_multiclass.to_uint8_masks(z, 0.8)
        test_multiclass.to_uint8_masks(z)
        subsets_union.return_union_sets_descending_order(g.clone())
        subsets_union.return_union_sets_descending_order(torch.sigmoid(z), reverse=True)
        train_multiclass.return_union_sets_descending_order(g.clone())
    # ragged / unaligned: odd H*W, strided channel slices
    z, g = torch.randn(3, 3, 17, 19).cuda(), (torch.rand(3, 3, 17, 19) > 0.5).float().cuda()
    zz = z.clone().requires_grad_(True)
    np.random.seed(0)
    combine(eco.losses_fn(torch.sigmoid(zz), g, True), UP).backward()
    zz = z.clone().requires_grad_(True)
    combine(train_multiclass.losses_fn(torch.sigmoid(zz), g, False, 0, False), UP).backward()
    test_multiclass.score_batch(z, g, list(np.arange(0.8, 0.99, 0.01)))
    test_multiclass.score_batch(z[:, 1:3], g[:, 1:3], 0.8)
    test_multiclass_sequential_densenetloss.score_batch(z, g)
    # frame pre-processing: odd sizes, up- and down-scaling
    for hin, win, out in ((90, 160, (32, 48)), (37, 53, (64, 64)), (270, 480, (128, 128))):
        fr = torch.from_numpy((np.random.RandomState(3).rand(2, hin, win, 3) * 255).astype(np.uint8)).cuda()
        test_video.preprocess_frames(fr, out)
    torch.cuda.synchronize()
    print("sanitize script ok")


if __name__ == "__main__":
    main()
