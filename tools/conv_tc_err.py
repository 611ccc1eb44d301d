"""Error of the tensor-core convolutions against float64, per operand format: maximum error relative to the output range and
the SIGNED mean relative error on large outputs (the tensor core truncates toward zero when it adds into its fp32 accumulator:
a systematic shrink that grows with the number of MMA steps per accumulator).  Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from side_b200 import ops
from side_b200.networks.stereo_network import cost_volume
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
cl = lambda x: x.permute(0, 2, 3, 4, 1).contiguous()
for fmt in ("tf32", "f16"):
    ops.set_tc_format(fmt)
    for (N, D, H, W, Cin, Cout, kd) in [(2, 16, 16, 16, 96, 64, 3), (2, 16, 8, 8, 128, 128, 3), (1, 16, 16, 16, 64, 128, 3),
                                        (1, 4, 48, 160, 64, 64, 1), (1, 4, 24, 80, 128, 128, 1), (1, 4, 24, 80, 256, 256, 1),
                                        (1, 8, 12, 40, 512, 512, 1)]:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(N, Cin, D, H, W, generator=g).relu()
        w = torch.randn(Cout, Cin, kd, 3, 3, generator=g) * (2.0 / (9 * kd * Cout)) ** 0.5
        ref = cl(F.conv3d(x.double(), w.double(), padding=(kd // 2, 1, 1)))
        hi, lo = ops.ncdhw_to_cl_split(x.to(dev), fmt=fmt)
        y, _, _ = ops.conv3d_tc(hi, lo, ops.conv_tc_prepare(w.to(dev), fmt=fmt), Cout, ksize=(kd, 3, 3), full=True, split=False)
        y = y[..., :Cout]
        e = (y.cpu().double() - ref)
        y32 = cl(F.conv3d(x.to(dev), w.to(dev), padding=(kd // 2, 1, 1))).cpu().double()
        e32 = y32 - ref
        big = ref.abs() > ref.abs().max() * 0.1
        steps = 9 * kd * Cin // (8 if fmt == "tf32" else 16)
        print("%s conv %s (%d MMA steps): max|err|/max|ref| tc %.2e cudnn-fp32 %.2e | signed mean rel err on large outputs tc %.2e "
              "(per step %.2e) cudnn %.2e" % (fmt, (N, D, H, W, Cin, Cout, kd), steps, e.abs().max() / ref.abs().max(),
                                            e32.abs().max() / ref.abs().max(), (e[big] / ref[big].abs() * ref[big].sign()).mean(),
                                            (e[big] / ref[big].abs() * ref[big].sign()).mean() / steps,
                                            (e32[big] / ref[big].abs() * ref[big].sign()).mean()), flush=True)
    for (N, D) in [(3, 16)]:
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
        print("%s aggregate N=%d D=%d: tc %.2e  cudnn-fp32 %.2e (max abs err / logit range)" % (fmt, N, D, (a - ref).abs().max() / s, (b - ref).abs().max() / s))
