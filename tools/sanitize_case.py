"""Small invocations of every kernel added this round (for compute-sanitizer --tool memcheck, one tool per GPU call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from side_b200 import ops
ops.set_tc_format("tf32")      # this tool feeds tf32 pairs (ops.tf32_split)
from side_b200.utils.synthetic import make_boxes
dev = torch.device("cuda")
torch.manual_seed(0)
# separable volume (incl. a box at the right border and an invalid row)
fL, fR = torch.randn(1, 8, 24, 320, device=dev), torch.randn(1, 8, 24, 320, device=dev)
left = torch.tensor([[0, 40.3, 3.2, 71.9, 17.5], [0, 300.0, 2.0, 318.5, 9.0], [0, 5.0, 1.0, 200.0, 20.0]], device=dev)
right = torch.tensor([[0, 33.1, 3.0, 65.2, 17.9], [0, 294.0, 2.5, 312.5, 9.5], [0, 1.0, 1.0, 190.0, 20.0]], device=dev)
fb = torch.tensor([384.38], device=dev)
valid = torch.tensor([1, 1, 1], dtype=torch.uint8, device=dev)
ops.inst_costvol_ungated(fL, fR, left, right, fb, 16, 16, 319.0, valid=valid)
ops.inst_costvol(fL, fR, left, right, fb, 16, 16, 319.0, gate=True, separable=True)
# tensor-core convolutions: role-swapped (Cout 64, 16x16), voxel-major (Cout 128, 8x8), 2-D strided, 1x1
x = torch.randn(2, 4, 16, 16, 32, device=dev); hi, lo = ops.tf32_split(x)
ops.conv3d_tc(hi, lo, ops.conv_tc_prepare(torch.randn(64, 32, 3, 3, 3, device=dev) * .1), 64, relu=True, full=True, split=True)
x = torch.randn(1, 4, 8, 8, 32, device=dev); hi, lo = ops.tf32_split(x)
ops.conv3d_tc(hi, lo, ops.conv_tc_prepare(torch.randn(128, 32, 3, 3, 3, device=dev) * .1), 128, relu=True, full=True, split=True)
x = torch.randn(1, 4, 16, 32, 32, device=dev); hi, lo = ops.tf32_split(x)
ops.conv3d_tc(hi, lo, ops.conv_tc_prepare(torch.randn(32, 32, 3, 3, device=dev) * .1), 32, ksize=(1, 3, 3), full=True, split=False, stride=2)
ops.conv3d_tc(hi, lo, ops.conv_tc_prepare(torch.randn(16, 32, 1, 1, device=dev) * .1), 16, ksize=(1, 1, 1), full=True, split=False)
# DCN channels-last path, stem, helpers, gwc
from side_b200.dcn_v2 import DCN
ops.set_dcn_precision("3xtf32")
m = DCN(32, 32, (3, 3), 1, 1).to(dev).eval()
with torch.no_grad():
    m.conv_offset_mask.weight.normal_(0, 0.05)
    m(torch.randn(4, 32, 16, 32, device=dev))
ops.stem_conv(torch.randn(1, 3, 20, 70, device=dev), torch.randn(16, 3, 7, 7, device=dev), stride=1)
ops.stem_conv(torch.randn(1, 16, 20, 70, device=dev), torch.randn(32, 16, 3, 3, device=dev), stride=2)
ops.gwc_volume(torch.randn(1, 16, 4, 64, device=dev), torch.randn(1, 16, 4, 64, device=dev), 8, 2)
ops.dw_deconv(torch.randn(1, 8, 6, 20, device=dev), torch.randn(8, 1, 4, 4, device=dev), 2, 1)
y = torch.randn(1, 2, 4, 4, 8, device=dev)
ops.maxpool_hw2_cl(y, full=True, split=True); ops.cl_to_nchw(y, 1, 8, (2, 4, 4))
torch.cuda.synchronize()
print("sanitize_case ok")
