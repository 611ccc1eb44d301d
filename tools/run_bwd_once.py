"""Two launches each of the backward kernels at their benchmark shapes (for ncu): the instance cost volume gather backward at
config #2 and the channels-last DCN backward of the 64 -> 64 @ 96x320 layer, B = 2."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from side_b200 import ops  # noqa: E402
from side_b200.utils.synthetic import make_boxes  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
fL, fR = torch.randn(1, 64, 96, 320, device=dev), torch.randn(1, 64, 96, 320, device=dev)
left, right, _ = make_boxes(1, 64, seed=0)
left, right, fb = left.to(dev), right.to(dev), torch.tensor([384.38], device=dev)
fLg, fRg = fL.clone().requires_grad_(True), fR.clone().requires_grad_(True)
cost, _ = ops.inst_costvol(fLg, fRg, left, right, fb, 48, 16, 319.0)
gcost = torch.randn_like(cost)
for _ in range(2):
    torch.autograd.grad(cost, (fLg, fRg), gcost, retain_graph=True)
del cost, gcost
x = torch.randn(2, 64, 96, 320, device=dev)
off = torch.randn(2, 18, 96, 320, device=dev) * 2
mask = torch.sigmoid(torch.randn(2, 9, 96, 320, device=dev))
w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
gy = torch.randn(2, 64, 96, 320, device=dev)
for _ in range(2):
    ops.dcn_backward_raw(x, off, mask, w, gy, 1, 1, 1, 1)
torch.cuda.synchronize()
print("ok")
