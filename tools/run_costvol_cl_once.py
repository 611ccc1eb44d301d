"""One-shot driver for ncu: the fused cost-volume builder (gated channels-last fp16 pairs) at the shape of one inference
micro-batch (800 RoIs x 16 candidates x 32 channels: 1.26 GB of pairs).
python tools/run_costvol_cl_once.py [N] [D] [wmin] [wmax]   box widths ~ U(wmin, wmax) feature columns (default 16..30: what the
network's wh head produces; 8..58 = BASELINE config #2's distribution, half of which needs the column-tiled two-pass mode)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from side_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 800
D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
wmin = float(sys.argv[3]) if len(sys.argv) > 3 else 16.0
wmax = float(sys.argv[4]) if len(sys.argv) > 4 else 30.0
B = 8
DIFF = os.environ.get("SIDE_CL_DIFF", "0") == "1"       # default: the (L, R)-only volume the network uses
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
fL, fR = torch.randn(B, 32, 96, 320, generator=g).to(dev), torch.randn(B, 32, 96, 320, generator=g).to(dev)
x1 = 20 + 230 * torch.rand(N, generator=g); w = wmin + (wmax - wmin) * torch.rand(N, generator=g)
y1 = 20 + 50 * torch.rand(N, generator=g); h = 6 + 25 * torch.rand(N, generator=g)
sh = 2 + 10 * torch.rand(N, generator=g)
b = torch.arange(B).repeat_interleave(N // B).float()
left = torch.stack([b, x1, y1, x1 + w, y1 + h], 1).to(dev)
right = torch.stack([b, x1 - sh, y1, x1 + w - sh, y1 + h], 1).to(dev)
fb = torch.full((B,), 384.38, device=dev)
for _ in range(4):
    hi, lo, db, xc = ops.inst_costvol_cl(fL, fR, left, right, fb, D, 16, 319.0, diff=DIFF)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.inst_costvol_cl(fL, fR, left, right, fb, D, 16, 319.0, diff=DIFF)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
byts = hi.numel() * 4 + 2 * fL.numel() * 4
print("inst_costvol_cl N=%d D=%d widths %.0f..%.0f: %.1f us per call (incl. the two NHWC staging copies), %.0f GB/s algorithmic" % (
    N, D, wmin, wmax, ms * 1000, byts / ms / 1e6))
