"""IDAUp step between its deformable convolutions: fused kernel vs dw_deconv + add + ncdhw_to_cl_split (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops
ops.set_tc_format("f16")
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
for B, C, H, W, f in [(32, 64, 48, 160, 2), (32, 64, 24, 80, 4), (32, 128, 24, 80, 2), (32, 256, 12, 40, 2)]:
    x = torch.randn(B, C, H, W, device=dev); skip = torch.randn(B, C, H * f, W * f, device=dev); w = torch.randn(C, 1, 2 * f, 2 * f, device=dev)
    a = t(lambda: ops.ncdhw_to_cl_split((ops.dw_deconv(x, w, f, f // 2) + skip).unsqueeze(2), want_full=True))
    b = t(lambda: ops.idaup_fuse_cl(x, w, skip, f))
    byts = (x.numel() + skip.numel() * 3) * 4
    print("B %d C %d %dx%d f%d: sequence %.3f ms, fused %.3f ms (%.0f GB/s)" % (B, C, H, W, f, a, b, byts / b / 1e6))
