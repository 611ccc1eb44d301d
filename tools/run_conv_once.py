"""A few launches of the tcgen05 conv3d at the aggregation network's layer shapes (for ncu).  argv: N D H W Cin Cout [tf32|f16 [mode]]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops
N, D, H, W, Cin, Cout = (int(v) for v in sys.argv[1:7])
fmt = sys.argv[7] if len(sys.argv) > 7 else "tf32"
ops.set_tc_format(fmt)
from side_b200 import _lib
_lib.load().side_conv_tc_set_mode(int(sys.argv[8]) if len(sys.argv) > 8 else 0)      # 32: no role-swapped kernel
dev = torch.device("cuda")
torch.manual_seed(0)
x = torch.randn(N, D, H, W, Cin, device=dev)
if fmt == "f16":
    hi = x.half()
    lo = ((x - hi.float()) * 2048.0).half()
else:
    hi, lo = ops.tf32_split(x)
w = torch.randn(Cout, Cin, 3, 3, 3, device=dev) * 0.05
wp = ops.conv_tc_prepare(w)
sc, sh = torch.rand(Cout, device=dev) + 0.5, torch.randn(Cout, device=dev)
for _ in range(3):
    ops.conv3d_tc(hi, lo, wp, Cout, scale=sc, shift=sh, relu=True, full=False, split=True)
torch.cuda.synchronize()
print("ok")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv3d_tc(hi, lo, wp, Cout, scale=sc, shift=sh, relu=True, full=False, split=True)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
print("%dx%dx%dx%d %d->%d %s: %.1f us  %.1f TF/s" % (N, D, H, W, Cin, Cout, fmt, us, 2.0 * N * D * H * W * Cin * Cout * 27 / us / 1e6))
