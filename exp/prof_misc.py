"""A few launches of the kernels added late in round 2 (for ncu): soft-label CE (128-bit kernels), soft Dice with the fused
prediction un-union, the plain 3-organ step on probabilities."""
import sys
import torch
sys.path.insert(0, ".")
import ecologysemanticsegmentation_b200 as eco
from ecologysemanticsegmentation_b200 import fused, ops
from ecologysemanticsegmentation_b200.synthetic import make_inputs
z, g = make_inputs(54, 3, 1024, 103)
z, g = z.cuda(), g.cuda()
for _ in range(3):
    ops.dice_counts_ex(z, g, None, ununion_preds=True)
    pr = torch.rand(54, 3, 1024, 1024, device="cuda").requires_grad_(True)
    eco.loss_functions.cross_entropy_loss(g, pr).backward()
z2, g2 = make_inputs(54, 3, 256, 102)
z2, g2 = z2.cuda(), g2.cuda()
up = torch.tensor(fused.loss_weights(bce=1.0, generalized_dice=1.0, twersky=1.0, focal_dice=1.0), dtype=torch.float32, device="cuda")
for _ in range(3):
    ops.multiclass3_fused(torch.sigmoid(z2), g2, 1.0, up, probs=True)
torch.cuda.synchronize()
print("ok")
