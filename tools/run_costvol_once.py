"""One launch of each instance cost-volume variant at config #2 (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops
from side_b200.utils.synthetic import make_boxes
dev = torch.device("cuda")
torch.manual_seed(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "separable"
B, C, N, D = 1, 64, 64, 48
fL, fR = torch.randn(B, C, 96, 320, device=dev), torch.randn(B, C, 96, 320, device=dev)
left, right, _ = make_boxes(B, N, seed=0)
left, right, fb = left.to(dev), right.to(dev), torch.full((B,), 384.38, device=dev)
for _ in range(3):
    if mode == "separable":
        ops.inst_costvol(fL, fR, left, right, fb, D, 16, 319.0, separable=True)
    elif mode == "ungated":
        ops.inst_costvol_ungated(fL, fR, left, right, fb, D, 16, 319.0)
    else:
        ops.inst_costvol(fL, fR, left, right, fb, D, 16, 319.0, gate=(mode == "gate"))
torch.cuda.synchronize()
print("ok")
