"""A few launches of the tcgen05 conv3d at the aggregation network's layer shapes (for ncu).  argv: N D H W Cin Cout"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops
ops.set_tc_format("tf32")      # this tool feeds tf32 pairs (ops.tf32_split)
N, D, H, W, Cin, Cout = (int(v) for v in sys.argv[1:7])
dev = torch.device("cuda")
torch.manual_seed(0)
x = torch.randn(N, D, H, W, Cin, device=dev)
hi, lo = ops.tf32_split(x)
w = torch.randn(Cout, Cin, 3, 3, 3, device=dev) * 0.05
wp = ops.conv_tc_prepare(w)
sc, sh = torch.rand(Cout, device=dev) + 0.5, torch.randn(Cout, device=dev)
for _ in range(3):
    ops.conv3d_tc(hi, lo, wp, Cout, scale=sc, shift=sh, relu=True, full=False, split=True)
torch.cuda.synchronize()
print("ok")
