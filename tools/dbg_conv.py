import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops, _lib
ops.set_tc_format("tf32")      # this tool feeds tf32 pairs (ops.tf32_split)
lib = _lib.load()
dev = torch.device("cuda")
for (N, D, H, W, Cin, Cout) in [(400, 16, 16, 16, 64, 64), (400, 16, 16, 16, 64, 128)]:
    x = torch.randn(N, D, H, W, Cin, device=dev); hi, lo = ops.tf32_split(x)
    wp = ops.conv_tc_prepare(torch.randn(Cout, Cin, 3, 3, 3, device=dev) * 0.05)
    for mode in (33, 32, 0):
        lib.side_conv_tc_set_mode(mode)
        for _ in range(2): ops.conv3d_tc(hi, lo, wp, Cout, relu=True, full=False, split=True)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): ops.conv3d_tc(hi, lo, wp, Cout, relu=True, full=False, split=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        fl = 2.0 * N * D * H * W * Cout * Cin * 27
        print(Cin, Cout, "mode", mode, "%.3f ms  %.0f TF/s" % (ms, fl / ms / 1e9))
lib.side_conv_tc_set_mode(0)
