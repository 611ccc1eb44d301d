"""A few launches of the tcgen05 convolution at a DLA 2-D layer shape (for ncu).  argv: B H W Cin Cout [k [mode]]  (fp16 pairs, full + split out;
mode = side_conv_tc_set_mode bit mask, 1 = halo box shared by the three vertical taps); prints the CUDA-event time per launch"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops
B, H, W, Cin, Cout = (int(v) for v in sys.argv[1:6])
k = int(sys.argv[6]) if len(sys.argv) > 6 else 3
ops.set_tc_format("f16")
from side_b200 import _lib
_lib.load().side_conv_tc_set_mode(int(sys.argv[7]) if len(sys.argv) > 7 else 0)
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.conv3d_tc(hi, lo, wp, Cout, ksize=(1, k, k), scale=sc, shift=sh, relu="after", residual=res, full=True, split=True)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print("B %d %dx%d %d->%d k%d mode %s: %.1f us  %.1f TF/s" % (B, H, W, Cin, Cout, k, sys.argv[7] if len(sys.argv) > 7 else "0", us, 2.0 * B * H * W * Cin * Cout * k * k / us / 1e6))
