"""A few launches of the tcgen05 convolution at a DLA 2-D layer shape (for ncu).  argv: B H W Cin Cout [k]  (fp16 pairs, full + split out)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops
B, H, W, Cin, Cout = (int(v) for v in sys.argv[1:6])
k = int(sys.argv[6]) if len(sys.argv) > 6 else 3
ops.set_tc_format("f16")
dev = torch.device("cuda")
torch.manual_seed(0)
x = torch.randn(1, B, H, W, Cin, device=dev)
hi = x.half(); lo = ((x - hi.float()) * 2048.0).half()
w = torch.randn(Cout, Cin, 1, k, k, device=dev) * 0.05
wp = ops.conv_tc_prepare(w)
sc, sh = torch.rand(Cout, device=dev) + 0.5, torch.randn(Cout, device=dev)
res = torch.randn(1, B, H, W, Cout, device=dev)
for _ in range(3):
    ops.conv3d_tc(hi, lo, wp, Cout, ksize=(1, k, k), scale=sc, shift=sh, relu="after", residual=res, full=True, split=True)
torch.cuda.synchronize()
print("ok")
