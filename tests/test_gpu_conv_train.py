"""GPU parity: training-mode convolutions on tcgen05 (side_b200/conv_train.py: forward, input gradient through the flipped
kernel, weight gradient through side_conv_wgrad_tc) against float64 autograd of torch's convolution -- the op the reference
trains with (nn.Conv2d / nn.Conv3d under stereoTrainer.py:254-319).  Bar: <= 1e-4 of each tensor's range."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x, w, gy, stride, pad):
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    conv = F.conv3d if x.dim() == 5 else F.conv2d
    y = conv(xd, wd, None, stride, pad)
    gx, gw = torch.autograd.grad(y, (xd, wd), gy.double())
    return y.detach(), gx, gw


def _err(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("cfg", [
    # (module, x shape, Cout, k, stride)
    ("2d", (4, 64, 96, 320), 64, 3, 1),         # DLA level 2 block
    ("2d", (4, 64, 96, 320), 128, 3, 2),        # stride-2 entry of level 3 (input gradient: cuDNN)
    ("2d", (4, 128, 48, 160), 128, 3, 1),
    ("2d", (4, 256, 24, 80), 256, 1, 1),        # Root 1x1
    ("2d", (4, 512, 12, 40), 512, 3, 1),        # 12 x 40: P = 480, the last 64-pixel k-block of a sample is partial
    ("2d", (2, 128, 96, 320), 256, 3, 1),       # stereo head input cat(L, R)
    ("3d", (16, 96, 16, 16, 16), 64, 3, 1),     # dres0.0: 96 channels, patch rows padded to 128
    ("3d", (16, 64, 16, 16, 16), 64, 3, 1),
    ("3d", (16, 128, 16, 8, 8), 128, 3, 1),
    ("3d", (16, 128, 16, 4, 4), 64, 3, 1),
])
def test_conv_train_forward_and_gradients(lib, cfg):
    from side_b200 import conv_train as ct
    kind, xs, Cout, k, stride = cfg
    torch.manual_seed(xs[1] + Cout)
    Cin = xs[1]
    m = (ct.TCConv2d if kind == "2d" else ct.TCConv3d)(Cin, Cout, k, stride=stride, padding=(k - 1) // 2, bias=False).cuda()
    x = torch.randn(xs, device="cuda").requires_grad_(True)
    ct._unsupported.clear()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False         # the pieces that stay on cuDNN run strict fp32, as in bench.py
    try:
        y = m(x)
        gy = torch.randn_like(y) * 1e-3             # a mean-reduced loss' scale
        gx, gw = torch.autograd.grad(y, (x, m.weight), gy)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    if stride == 1:
        assert not ct._unsupported, ct._unsupported  # forward, input gradient and weight gradient all ran on tcgen05
    ry, rgx, rgw = _ref(x.detach(), m.weight.detach(), gy, stride, (k - 1) // 2)
    assert ("fwd", (tuple(x.unsqueeze(2).shape) if kind == "2d" else tuple(xs), Cout, ((1, k, k) if kind == "2d" else (k, k, k)), stride)) \
        not in ct._unsupported, "forward fell back to cuDNN"
    assert _err(y, ry) < 1e-4 and _err(gx, rgx) < 1e-4 and _err(gw, rgw) < 1e-4, (_err(y, ry), _err(gx, rgx), _err(gw, rgw))
    # the same module against cuDNN fp32 (what the reference computes): both within the bar of the float64 result
    ct.enabled = False
    torch.backends.cudnn.allow_tf32 = False         # strict fp32 cuDNN (its default lets TF32 in: ~3e-4 on these sums)
    try:
        y2 = m(x)
        gx2, gw2 = torch.autograd.grad(y2, (x, m.weight), gy)
    finally:
        ct.enabled = True
        torch.backends.cudnn.allow_tf32 = old
    assert _err(gw, gw2.double()) < 2e-4 and _err(gx, gx2.double()) < 2e-4


def test_conv_train_gradient_scales(lib):
    """grad_output of arbitrary magnitude (1e-9 .. 1e+5): the power-of-two range scale keeps the fp16 pairs exact enough."""
    from side_b200 import conv_train as ct
    torch.manual_seed(0)
    m = ct.TCConv3d(64, 64, 3, padding=1, bias=False).cuda()
    x = torch.randn(4, 64, 16, 16, 16, device="cuda").requires_grad_(True)
    y = m(x)
    for scale in (1e-9, 1e-4, 1.0, 1e5):
        gy = torch.randn_like(y) * scale
        gx, gw = torch.autograd.grad(y, (x, m.weight), gy, retain_graph=True)
        _, rgx, rgw = _ref(x.detach(), m.weight.detach(), gy, 1, 1)
        assert _err(gx, rgx) < 1e-4 and _err(gw, rgw) < 1e-4, scale


def test_eval_mode_is_plain_conv(lib):
    from side_b200 import conv_train as ct
    m = ct.TCConv2d(64, 64, 3, padding=1, bias=True).cuda().eval()
    x = torch.randn(2, 64, 24, 80, device="cuda")
    with torch.no_grad():
        assert torch.equal(m(x), F.conv2d(x, m.weight, m.bias, 1, 1))
    y = m(x.requires_grad_(True))                     # autograd on: tensor-core path, bias added afterwards
    ref = F.conv2d(x.double(), m.weight.double(), m.bias.double(), 1, 1)
    assert _err(y.detach(), ref.detach()) < 1e-4


@pytest.mark.parametrize("shape", [(4, 64, 96, 320), (4, 512, 12, 40), (16, 128, 16, 8, 8), (16, 64, 16, 16, 16), (3, 32, 7, 9)])
def test_batchnorm_train_matches_torch(lib, shape):
    """TCBatchNorm (side_bn_train_fwd / _bwd) against nn.BatchNorm in float64: output, running statistics, all three gradients."""
    from side_b200 import conv_train as ct
    torch.manual_seed(shape[1])
    C = shape[1]
    cls, ref_cls = (ct.TCBatchNorm2d, torch.nn.BatchNorm2d) if len(shape) == 4 else (ct.TCBatchNorm3d, torch.nn.BatchNorm3d)
    m = cls(C, momentum=0.1).cuda().train()
    ref = ref_cls(C, momentum=0.1).cuda().double().train()
    with torch.no_grad():
        m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
        ref.weight.copy_(m.weight); ref.bias.copy_(m.bias)
    x = (torch.randn(shape, device="cuda") * 2 + 0.7).requires_grad_(True)
    xd = x.detach().double().requires_grad_(True)
    y = m(x)
    yr = ref(xd)
    gy = torch.randn_like(y)
    gx, gw, gb = torch.autograd.grad(y, (x, m.weight, m.bias), gy)
    rgx, rgw, rgb = torch.autograd.grad(yr, (xd, ref.weight, ref.bias), gy.double())
    for a, b, name in ((y, yr, "y"), (gx, rgx, "gx"), (gw, rgw, "gweight"), (gb, rgb, "gbias"),
                       (m.running_mean, ref.running_mean, "running_mean"), (m.running_var, ref.running_var, "running_var")):
        assert _err(a.detach(), b.detach()) < 1e-5, name
    assert int(m.num_batches_tracked) == 1
    m.eval()
    with torch.no_grad():
        assert _err(m(x), ref.eval()(xd)) < 1e-5
