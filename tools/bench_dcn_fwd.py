"""DCN forward sweep (BASELINE config #3 shapes, B = 2 and 16) for the tensor-core precisions vs torchvision's CUDA op."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from side_b200 import ops
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def time_op(fn, iters=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(iters):
        flush.zero_(); torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / iters
shapes = [(512, 256, 12, 40), (256, 256, 24, 80), (256, 128, 24, 80), (128, 128, 48, 160), (128, 64, 48, 160), (64, 64, 96, 320), (256, 64, 24, 80)]
for B in (2, 16):
    for (Cin, Cout, H, W) in shapes:
        torch.manual_seed(0)
        x = torch.randn(B, Cin, H, W, device=dev); off = torch.randn(B, 18, H, W, device=dev) * 2
        mask = torch.sigmoid(torch.randn(B, 9, H, W, device=dev)); w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05
        b = torch.rand(Cout, device=dev)
        r = {p: time_op(lambda: ops.dcn_forward_raw(x, off, mask, w, b, 1, 1, 1, 1, precision=p)) for p in ("3xtf32", "3xfp16")}
        tv = time_op(lambda: torchvision.ops.deform_conv2d(x, off, w, b, padding=1, mask=mask))
        ref = ops.dcn_forward_raw(x, off, mask, w, b, 1, 1, 1, 1, precision="fp32")
        err = {p: float((ops.dcn_forward_raw(x, off, mask, w, b, 1, 1, 1, 1, precision=p) - ref).abs().max() / ref.abs().max()) for p in r}
        print("B=%2d %3dx%3d@%2dx%3d  3xtf32 %.4f ms  3xfp16 %.4f ms  torchvision %.4f ms  (x%.2f / x%.2f)  err %.1e / %.1e" % (
            B, Cin, Cout, H, W, r["3xtf32"], r["3xfp16"], tv, tv / r["3xtf32"], tv / r["3xfp16"], err["3xtf32"], err["3xfp16"]))
