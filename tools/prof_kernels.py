"""Runs each hot kernel a few times at its BASELINE shape (for `ncu --set full -k regex:...` captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops
from side_b200.utils.synthetic import make_boxes

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda")
torch.manual_seed(0)
fL, fR = torch.randn(1, 64, 96, 320, device=dev), torch.randn(1, 64, 96, 320, device=dev)
left, right, _ = make_boxes(1, 64, seed=0)
left, right, fb = left.to(dev), right.to(dev), torch.tensor([384.38], device=dev)
for it in range(3):
    if which in ("all", "costvol"):
        ops.inst_costvol(fL, fR, left, right, fb, 48, 16, 319.0, gate=True)
    if which in ("all", "costvol_nchw"):
        ops.USE_NHWC_GATHER = False
        ops.inst_costvol(fL, fR, left, right, fb, 48, 16, 319.0, gate=True)
        ops.USE_NHWC_GATHER = True
    if which in ("all", "concat"):
        ops.concat_volume(fL, fR, 48)
    if which in ("all", "dcn"):
        x = torch.randn(2, 64, 96, 320, device=dev); off = torch.randn(2, 18, 96, 320, device=dev) * 2
        m = torch.sigmoid(torch.randn(2, 9, 96, 320, device=dev)); w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
        b = torch.rand(64, device=dev)
        ops.dcn_forward_raw(x, off, m, w, b, 1, 1, 1, 1, precision="3xtf32")
        x = torch.randn(2, 256, 24, 80, device=dev); off = torch.randn(2, 18, 24, 80, device=dev) * 2
        m = torch.sigmoid(torch.randn(2, 9, 24, 80, device=dev)); w = torch.randn(256, 256, 3, 3, device=dev) * 0.05
        b = torch.rand(256, device=dev)
        ops.dcn_forward_raw(x, off, m, w, b, 1, 1, 1, 1, precision="3xtf32")
    if which in ("all", "decode"):
        hm = torch.randn(4, 3, 96, 320, device=dev); wh = torch.rand(4, 3, 96, 320, device=dev); reg = torch.rand(4, 3, 96, 320, device=dev)
        ops.bbox_decode_raw(hm, wh, reg, K=100)
torch.cuda.synchronize()
print("ok")
