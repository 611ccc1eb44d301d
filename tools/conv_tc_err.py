"""Prints the error of the tensor-core aggregation path against float64 (per conv and end to end)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from side_b200 import ops
ops.set_tc_format("tf32")      # this tool feeds tf32 pairs (ops.tf32_split)
from side_b200.networks.stereo_network import cost_volume
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
cl = lambda x: x.permute(0, 2, 3, 4, 1).contiguous()
for (N, D, H, W, Cin, Cout) in [(2, 16, 16, 16, 96, 64), (2, 16, 8, 8, 128, 128), (1, 16, 16, 16, 64, 128)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, Cin, D, H, W, generator=g).relu()
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) * (2.0 / (27 * Cout)) ** 0.5
    ref = cl(F.conv3d(x.double(), w.double(), padding=1))
    hi, lo = ops.tf32_split(cl(x).to(dev))
    y, _, _ = ops.conv3d_tc(hi, lo, ops.conv_tc_prepare(w.to(dev)), Cout, full=True, split=False)
    e = (y.cpu().double() - ref)
    y32 = cl(F.conv3d(x.to(dev), w.to(dev), padding=1)).cpu().double()
    e32 = y32 - ref
    big = ref.abs() > ref.abs().max() * 0.1
    print("conv %s: max|err|/max|ref| tc %.2e cudnn-fp32 %.2e | mean signed rel err on large outputs tc %.2e cudnn %.2e" % (
        (N, D, H, W, Cin, Cout), e.abs().max() / ref.abs().max(), e32.abs().max() / ref.abs().max(),
        (e[big] / ref[big].abs() * ref[big].sign()).mean(), (e32[big] / ref[big].abs() * ref[big].sign()).mean()))
for (N, D) in [(3, 16), (2, 48)]:
    torch.manual_seed(11)
    m = cost_volume(64).eval()
    for mod in m.modules():
        if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.8, 1.2); mod.bias.data.normal_(0, 0.1)
    cost = torch.randn(N, 96, D, 16, 16)
    with torch.no_grad():
        ref = m.double().aggregate(cost.double())
        m = m.float().to(dev)
        a = m.aggregate_tc(cost.to(dev)).cpu().double()
        b = m.aggregate(cost.to(dev)).cpu().double()
    s = ref.abs().max()
    print("aggregate N=%d D=%d: tc %.2e  cudnn-fp32 %.2e (max abs err / logit range)" % (N, D, (a - ref).abs().max() / s, (b - ref).abs().max() / s))
