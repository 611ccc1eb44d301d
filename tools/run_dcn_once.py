"""A few launches of the tcgen05 DCN forward at one DLA-34 up-path shape (for ncu).  argv: Cin Cout H W B [zero]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops
Cin, Cout, H, W, B = (int(v) for v in sys.argv[1:6])
zero = len(sys.argv) > 6
dev = torch.device("cuda")
torch.manual_seed(0)
x = torch.randn(B, Cin, H, W, device=dev)
off = torch.zeros(B, 18, H, W, device=dev) if zero else torch.randn(B, 18, H, W, device=dev) * 2
mask = torch.sigmoid(torch.randn(B, 9, H, W, device=dev)); w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
b = torch.rand(Cout, device=dev)
for _ in range(3):
    ops.dcn_forward_raw(x, off, mask, w, b, 1, 1, 1, 1, precision=os.environ.get("SIDE_DCN_PRECISION", "3xtf32"))
torch.cuda.synchronize()
print("ok")
