"""Separable volume forward at config #2 / network / reference shapes (experiments: SIDE_SEP_Z, SIDE_SEP_NOSPLIT)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops
from side_b200.utils.synthetic import make_boxes
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def time_op(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / iters
torch.manual_seed(0)
tag = "Z=%s nosplit=%s" % (os.environ.get("SIDE_SEP_Z"), os.environ.get("SIDE_SEP_NOSPLIT"))
for name, (B, C, N, D) in {"config2": (1, 64, 64, 48), "network": (4, 32, 400, 16), "refshape": (1, 32, 100, 16)}.items():
    fL, fR = torch.randn(B, C, 96, 320, device=dev), torch.randn(B, C, 96, 320, device=dev)
    left, right, _ = make_boxes(B, N // B, seed=0)
    left, right, fb = left.to(dev), right.to(dev), torch.full((B,), 384.38, device=dev)
    byts = N * 3 * C * D * 256 * 4 + 2 * B * C * 96 * 320 * 4
    for label, fn in [("separable", lambda: ops.inst_costvol(fL, fR, left, right, fb, D, 16, 319.0, separable=True)),
                      ("ungated+xcross", lambda: ops.inst_costvol_ungated(fL, fR, left, right, fb, D, 16, 319.0))]:
        ms = time_op(fn)
        print("%-22s %-8s %-16s %.4f ms  %7.1f GB/s  %.3f" % (tag, name, label, ms, byts / ms / 1e6, byts / ms / 1e6 / 6460.2))
