import sys, torch
sys.path.insert(0, ".")
from ecologysemanticsegmentation_b200 import ops
from ecologysemanticsegmentation_b200.synthetic import make_inputs
z, g = make_inputs(54, 3, 512, 104)
p, g = torch.sigmoid(z.cuda()), g.cuda()
up = torch.tensor([0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 1.0], dtype=torch.float32, device="cuda")
for _ in range(3):
    sums = ops.pair_stats(g, p, 0)
    _, _, jac = ops.pair_finalize(sums, 0.0, [1.0] * 3)
    ops.pair_grad(g, p, 0, jac, up, False, True)
torch.cuda.synchronize()
print("ok")
