"""GPU parity: the aggregation network on the tensor cores.

Reference = the reference module's own ops (``nn.Conv3d`` / ``BatchNorm3d`` / ``MaxPool3d``, cost_volume.forward,
stereo_network_old.py:205-227) evaluated by PyTorch on the CPU in float64 -- a floating-point kernel, so the checker is
the plain torch op; tolerance 1e-4 relative (north_star: fp32 correlation / depth outputs)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def strict_fp32():
    """The strAM Conv2d stays cuDNN: keep it in true fp32 (torch lets cuDNN use TF32 by default)."""
    from side_b200 import ops
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    old_fmt = ops.get_tc_format()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ops.set_tc_format("tf32")          # the tests of this file name the operand format they exercise (tc_fmt fixture / fmt=...)
    yield
    ops.set_tc_format(old_fmt)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.fixture(params=["tf32", "f16"])
def tc_fmt(request):
    """Run the test under both operand formats of the tensor-core convolutions (3xTF32 / 3xFP16)."""
    from side_b200 import ops
    ops.set_tc_format(request.param)
    yield request.param


def _cl(x):      # NCDHW -> NDHWC
    return x.permute(0, 2, 3, 4, 1).contiguous()


@pytest.mark.parametrize("cfg", [
    # N, D, H, W, Cin, Cout, relu, affine, residual
    (2, 16, 16, 16, 96, 64, True, True, False),      # dres0.0
    (1, 8, 16, 16, 64, 128, True, True, False),      # dres1.3
    (3, 16, 8, 8, 128, 128, True, True, True),       # dres2.3 + residual
    (2, 16, 4, 4, 128, 64, True, True, False),       # classify.0
    (1, 48, 16, 16, 32, 16, False, False, False),    # plain conv, D = 48
    (5, 2, 8, 8, 64, 32, False, True, False),        # odd sample count, persistent tail
])
@pytest.mark.parametrize("mode", [0, 1, 32])
def test_conv3d_tc_matches_fp64(lib, cfg, mode):
    """mode 0: default (role-swapped kernel for Cout = 64 on 16x16 maps); 1: halo-box reuse in the voxel-major kernel;
    32: voxel-major kernel everywhere."""
    from side_b200 import ops
    lib.side_conv_tc_set_mode(mode)
    N, D, H, W, Cin, Cout, relu, affine, res = cfg
    g = torch.Generator().manual_seed(N * 1000 + Cin + Cout)
    x = torch.randn(N, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) * (2.0 / (27 * Cout)) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5 if affine else None
    shift = torch.randn(Cout, generator=g) * 0.1 if affine else None
    r = torch.randn(N, Cout, D, H, W, generator=g) if res else None
    ref = F.conv3d(x.double(), w.double(), padding=1)
    if affine:
        ref = ref * scale.double().view(1, -1, 1, 1, 1) + shift.double().view(1, -1, 1, 1, 1)
    if relu:
        ref = ref.relu()
    if res:
        ref = ref + r.double()
    ref = _cl(ref).numpy()

    dev = torch.device("cuda")
    wp = ops.conv_tc_prepare(w.to(dev))
    hi, lo = ops.tf32_split(_cl(x).to(dev))
    assert torch.equal(hi + lo, _cl(x).to(dev))                       # exact split
    y, yh, yl = ops.conv3d_tc(hi, lo, wp, Cout, scale=None if scale is None else scale.to(dev),
                              shift=None if shift is None else shift.to(dev), relu=relu,
                              residual=None if r is None else _cl(r).to(dev), full=True, split=True)
    assert rel_err(y.cpu().numpy(), ref) < 1e-4
    assert torch.equal(yh + yl, y)                                    # split outputs reassemble the fp32 result exactly
    assert torch.equal(yh, (y.view(torch.int32) & -8192).view(torch.float32))
    lib.side_conv_tc_set_mode(0)


def test_layout_and_pool_helpers(lib):
    from side_b200 import ops
    g = torch.Generator().manual_seed(3)
    dev = torch.device("cuda")
    x = torch.randn(3, 40, 4, 6, 10, generator=g)                    # NCDHW, ragged vs the 32x32 transpose tile
    hi, lo = ops.ncdhw_to_cl_split(x.to(dev))
    assert torch.equal((hi + lo).cpu(), _cl(x))
    assert torch.equal(hi.cpu(), (_cl(x).view(torch.int32) & -8192).view(torch.float32))
    sc = torch.rand(3, 4, generator=g) + 0.5                          # per-(sample, depth slice) scale (deferred cosine gate)
    hi, lo = ops.ncdhw_to_cl_split(x.to(dev), scale=sc.to(dev))
    assert torch.equal((hi + lo).cpu(), _cl(x * sc[:, None, :, None, None]))
    y = torch.randn(2, 5, 6, 8, 12, generator=g)                     # NDHWC
    gate = torch.rand(2, 5, 8, 12, generator=g)
    hi, lo = ops.gate_mul_split(y.to(dev), gate.to(dev))
    assert torch.equal((hi + lo).cpu(), y * gate.unsqueeze(2))
    full, hi, lo = ops.maxpool_hw2_cl(y.to(dev), full=True, split=True)
    ref = F.max_pool3d(y.permute(0, 4, 1, 2, 3), (1, 2, 2)).permute(0, 2, 3, 4, 1)
    assert torch.equal(full.cpu(), ref) and torch.equal((hi + lo).cpu(), ref)
    w = torch.randn(1, 12, 3, 3, 3, generator=g)
    out = ops.conv3d_c1_cl(y.to(dev), w.to(dev))
    ref = F.conv3d(y.permute(0, 4, 1, 2, 3).double(), w.double(), padding=1)[:, 0]
    assert rel_err(out.cpu().numpy(), ref.numpy()) < 1e-5


@pytest.mark.parametrize("N,D", [(3, 16), (2, 48)])
def test_aggregate_tc_matches_module_fp64(lib, N, D):
    """cost_volume.aggregate on tcgen05 == the reference layer sequence in float64 on the CPU (<= 1e-4 of the logit range),
    and the depth regressed from it agrees to 1e-4 relative."""
    from side_b200 import ops
    from side_b200.networks.stereo_network import cost_volume
    torch.manual_seed(11)
    m = cost_volume(64).eval()
    for mod in m.modules():                                           # non-trivial eval BatchNorm statistics
        if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.8, 1.2)
            mod.bias.data.normal_(0, 0.1)
    cost = torch.randn(N, 96, D, 16, 16)
    with torch.no_grad():
        ref = m.double().aggregate(cost.double()).float()
    m = m.float().cuda()
    with torch.no_grad():
        c = cost.cuda()
        assert m._tc_ok(c)
        out = m.aggregate_tc(c)
        assert tuple(out.shape) == (N, D, 4, 4)
        scale = float(ref.abs().max())
        assert float((out.cpu() - ref).abs().max()) < 1e-4 * scale
        db = torch.linspace(87, 5, D).repeat(N, 1)
        d_tc = ops.softargmin(out.contiguous(), db.cuda()).cpu()
        d_ref = ops.softargmin(ref.cuda().contiguous(), db.cuda()).cpu()
        assert float(((d_tc - d_ref).abs() / d_ref.abs()).max()) < 1e-4
        # the module switches back to cuDNN under autograd / training
        assert not m.train()._tc_ok(c)


@pytest.mark.parametrize("shape", [(2, 96, 320), (1, 16, 320), (3, 8, 64)])
def test_heads_tc_match_cudnn_heads(lib, shape, tc_fmt):
    """stereo_network heads (:343-348) on tcgen05 (stacked first convolutions, n-tiles of 128, 2 x 64 pixel boxes) vs the
    module's own cuDNN fp32 path on the same weights: <= 1e-4 of each head's range."""
    from side_b200.networks import get_pose_net
    from side_b200.utils.synthetic import HEADS, realistic_init
    B, H, W = shape
    torch.manual_seed(2)
    m = realistic_init(get_pose_net(34, HEADS, 256), seed=3).eval().cuda()
    fl, fr = torch.randn(B, 64, H, W, device="cuda"), torch.randn(B, 64, H, W, device="cuda")
    with torch.no_grad():
        assert m._heads_tc_ok(fl)
        z = m._heads_tc(fl, fr)
        both = torch.cat((fl, fr), 1)
        for h in m.heads:
            ref = m.__getattr__(h)(fl if h in m.left_only else both)
            assert z[h].shape == ref.shape and z[h].is_contiguous()
            err = float((z[h] - ref).abs().max() / ref.abs().max())
            assert err < 1e-4, (h, err)


def test_conv2d_tc_wide_n_tiles(lib):
    """2-D convolution through the same kernel: D = 1, kd = 1, W = 320 (boxes of 2 rows x 64 columns), Cout = 256 (2 n-tiles)."""
    from side_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 8, 320, generator=g)
    w = torch.randn(256, 64, 3, 3, generator=g) * 0.05
    ref = F.conv2d(x.double(), w.double(), padding=1).relu().permute(0, 2, 3, 1).numpy()
    dev = torch.device("cuda")
    hi, lo = ops.ncdhw_to_cl_split(x.to(dev).unsqueeze(2))
    y, yh, yl = ops.conv3d_tc(hi, lo, ops.conv_tc_prepare(w.to(dev)), 256, ksize=(1, 3, 3), relu=True, full=True, split=True)
    assert rel_err(y.cpu().numpy()[:, 0], ref) < 1e-4
    assert torch.equal(yh + yl, y)


def test_conv2d_tc_stride2_and_1x1(lib):
    """Stride-2 3x3 (TMA element strides) and 1x1 kernels with the batch as the box's depth axis; ReLU after the residual."""
    from side_b200 import ops
    g = torch.Generator().manual_seed(9)
    dev = torch.device("cuda")
    x = torch.randn(4, 32, 24, 80, generator=g)
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous().unsqueeze(0)        # [1, B, H, W, C]
    hi, lo = ops.tf32_split(cl(x).to(dev))
    w = torch.randn(64, 32, 3, 3, generator=g) * 0.1
    ref = cl(F.conv2d(x.double(), w.double(), stride=2, padding=1)).numpy()
    y, _, _ = ops.conv3d_tc(hi, lo, ops.conv_tc_prepare(w.to(dev)), 64, ksize=(1, 3, 3), full=True, split=False, stride=2)
    assert tuple(y.shape) == (1, 4, 12, 40, 64)
    assert rel_err(y.cpu().numpy(), ref) < 1e-4
    w1 = torch.randn(48, 32, 1, 1, generator=g) * 0.2
    res = torch.randn(4, 48, 24, 80, generator=g)
    ref = cl((F.conv2d(x.double(), w1.double()) + res.double()).relu()).numpy()
    y, _, _ = ops.conv3d_tc(hi, lo, ops.conv_tc_prepare(w1.to(dev)), 48, ksize=(1, 1, 1), relu="after", residual=cl(res).to(dev),
                            full=True, split=False)
    assert rel_err(y.cpu().numpy(), ref) < 1e-4


@pytest.mark.parametrize("B,H,W", [(4, 192, 640), (8, 64, 128), (2, 192, 640), (3, 64, 128)])
def test_dla_levels_tc_match_cudnn(lib, B, H, W, tc_fmt):
    """DLA-34 levels 2-5 on tcgen05 vs the same modules on cuDNN fp32: every level output <= 1e-4 of its range."""
    from side_b200.networks.feature_extraction_dla34 import dla34
    torch.manual_seed(4)
    m = dla34().eval()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.8, 1.2); mod.bias.data.normal_(0, 0.1)
    m = m.cuda()
    x = torch.randn(B, 32, H, W, device="cuda")          # level-1 output
    with torch.no_grad():
        assert m._tc_ok(x)                               # B = 2 / 3: one stereo pair (the detector's real batch) -- the depth boxes
        outs = m._levels_tc(x)                           # of the 12x40 / 4x8 levels hang over the batch (TMA zero fill)
        ref, t = [], x
        for i in range(2, 6):
            t = getattr(m, "level%d" % i)(t)
            ref.append(t)
    for i, (a, b) in enumerate(zip(outs, ref)):
        assert a.shape == b.shape
        err = float((a - b).abs().max() / b.abs().max())
        assert err < 1e-4, (i + 2, err)


@pytest.mark.parametrize("cfg", [(3, 16, 7, 1, 2, 70, 150), (16, 16, 3, 1, 2, 33, 129), (16, 32, 3, 2, 3, 64, 130), (16, 32, 3, 2, 1, 384, 1280)])
def test_stem_conv_matches_fp64(lib, cfg):
    """DLA-34 stem layers as direct fp32 convolutions with folded BatchNorm + ReLU (ragged tiles, odd widths, stride 2)."""
    from side_b200 import ops
    Cin, Cout, k, s, B, H, W = cfg
    g = torch.Generator().manual_seed(Cin + k + W)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    scale, shift = torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g) * 0.1
    ref = (F.conv2d(x.double(), w.double(), stride=s, padding=(k - 1) // 2) * scale.double().view(1, -1, 1, 1)
           + shift.double().view(1, -1, 1, 1)).relu()
    y = ops.stem_conv(x.cuda(), w.cuda(), scale.cuda(), shift.cuda(), stride=s, relu=True)
    assert y.shape == ref.shape
    assert rel_err(y.cpu().numpy(), ref.numpy()) < 1e-5


def test_dla_base_fast_paths_match_cudnn(lib, tc_fmt):
    """Whole DLA-34 base: direct stem + tcgen05 levels 2-5 vs the plain module on cuDNN fp32."""
    from side_b200.networks.feature_extraction_dla34 import dla34
    torch.manual_seed(6)
    m = dla34().eval()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.8, 1.2); mod.bias.data.normal_(0, 0.1)
    m = m.cuda()
    x = torch.randn(4, 3, 128, 256, device="cuda")
    with torch.no_grad():
        fast = m(x)
        m.direct_stem = m.tensor_core = False
        ref = m(x)
    assert len(fast) == len(ref) == 6
    for i, (a, b) in enumerate(zip(fast, ref)):
        assert a.shape == b.shape
        assert float((a - b).abs().max() / b.abs().max()) < 1e-4, i


# ------------------------------------------------------------------------------------------------
# "3xFP16" operand format: kind::f16 MMAs on fp16 pairs (hi, lo * 2^11), same 22 significand bits as the tf32 pair
# ------------------------------------------------------------------------------------------------
def _f16_split(x):
    hi = x.half()
    lo = ((x - hi.float()) * 2048.0).half()
    return hi, lo


@pytest.mark.parametrize("cfg", [
    # N, D, H, W, Cin, Cout, relu, affine, residual
    (2, 16, 16, 16, 64, 64, True, True, False),      # dres0.3 / dres1.0: role-swapped kernel
    (1, 8, 16, 16, 192, 64, True, True, False),      # three k-blocks per tap, role-swapped kernel
    (1, 8, 16, 16, 64, 128, True, True, False),      # dres1.3
    (3, 16, 8, 8, 128, 128, True, True, True),       # dres2.3 + residual
    (2, 16, 4, 4, 128, 64, True, True, False),       # classify.0
    (5, 2, 8, 8, 64, 32, False, False, False),       # odd sample count, no affine
    (1, 8, 16, 16, 96, 64, True, True, False),       # dres0.0: half-full last channel block (TMA zero fill, MMAs skipped)
    (2, 4, 8, 8, 32, 32, False, True, False),        # single half-full block
    (1, 8, 16, 16, 160, 128, True, False, False),    # 2.5 blocks, voxel-major kernel
])
@pytest.mark.parametrize("mode", [0, 32])
def test_conv3d_tc_f16_matches_fp64(lib, cfg, mode):
    from side_b200 import ops
    lib.side_conv_tc_set_mode(mode)
    N, D, H, W, Cin, Cout, relu, affine, res = cfg
    g = torch.Generator().manual_seed(N * 1000 + Cin + Cout + 7)
    x = torch.randn(N, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) * (2.0 / (27 * Cout)) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5 if affine else None
    shift = torch.randn(Cout, generator=g) * 0.1 if affine else None
    r = torch.randn(N, Cout, D, H, W, generator=g) if res else None
    ref = F.conv3d(x.double(), w.double(), padding=1)
    if affine:
        ref = ref * scale.double().view(1, -1, 1, 1, 1) + shift.double().view(1, -1, 1, 1, 1)
    if relu:
        ref = ref.relu()
    if res:
        ref = ref + r.double()
    ref = _cl(ref).numpy()
    dev = torch.device("cuda")
    wp = ops.conv_tc_prepare(w.to(dev), fmt="f16")
    hi, lo = _f16_split(_cl(x).to(dev))
    y, yh, yl = ops.conv3d_tc(hi, lo, wp, Cout, scale=None if scale is None else scale.to(dev),
                              shift=None if shift is None else shift.to(dev), relu=relu,
                              residual=None if r is None else _cl(r).to(dev), full=True, split=True)
    lib.side_conv_tc_set_mode(0)
    assert rel_err(y.cpu().numpy(), ref) < 1e-4
    assert yh.dtype == torch.float16 and torch.equal(yh, y.half())
    rec = yh.float() + yl.float() / 2048.0                           # the pair carries the fp32 result to ~2^-21
    assert float((rec - y).abs().max()) <= 2e-6 * float(y.abs().max())


def test_f16_layout_and_pool_helpers(lib):
    from side_b200 import ops
    g = torch.Generator().manual_seed(4)
    dev = torch.device("cuda")
    x = torch.randn(3, 40, 4, 6, 10, generator=g)                    # 40 channels -> padded to 64
    sc = torch.rand(3, 4, generator=g) + 0.5
    hi, lo = ops.ncdhw_to_cl_split(x.to(dev), scale=sc.to(dev), fmt="f16")
    assert tuple(hi.shape) == (3, 4, 6, 10, 64) and hi.dtype == torch.float16
    want = _cl(x * sc[:, None, :, None, None])
    rec = (hi.float() + lo.float() / 2048.0).cpu()
    assert float((rec[..., :40] - want).abs().max()) <= 2e-6 * float(want.abs().max())
    assert float(rec[..., 40:].abs().max()) == 0.0
    y = torch.randn(2, 5, 6, 8, 12, generator=g)
    gate = torch.rand(2, 5, 8, 12, generator=g)
    hi, lo = ops.gate_mul_split(y.to(dev), gate.to(dev), fmt="f16")
    want = y * gate.unsqueeze(2)
    assert float(((hi.float() + lo.float() / 2048.0).cpu() - want).abs().max()) <= 2e-6 * float(want.abs().max())
    full, hi, lo = ops.maxpool_hw2_cl(y.to(dev), full=True, split=True, fmt="f16")
    ref = F.max_pool3d(y.permute(0, 4, 1, 2, 3), (1, 2, 2)).permute(0, 2, 3, 4, 1)
    assert torch.equal(full.cpu(), ref)
    assert float(((hi.float() + lo.float() / 2048.0).cpu() - ref).abs().max()) <= 2e-6 * float(ref.abs().max())


@pytest.mark.parametrize("N,D", [(3, 16), (2, 48)])
def test_aggregate_tc_f16_matches_module_fp64(lib, N, D):
    """The aggregation network with tc_format = "f16" against the float64 layer sequence: same bars as the tf32 path."""
    from side_b200 import ops
    from side_b200.networks.stereo_network import cost_volume
    torch.manual_seed(11)
    m = cost_volume(64).eval()
    for mod in m.modules():
        if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.8, 1.2)
            mod.bias.data.normal_(0, 0.1)
    cost = torch.randn(N, 96, D, 16, 16)
    xc = torch.rand(N, D) * 0.8 + 0.1
    with torch.no_grad():
        ref = m.double().aggregate((cost * xc[:, None, :, None, None]).double()).float()
    m = m.float().cuda()
    m.tc_format = "f16"
    with torch.no_grad():
        out = m.aggregate_tc(cost.cuda(), xcross=xc.cuda())
        scale = float(ref.abs().max())
        assert float((out.cpu() - ref).abs().max()) < 1e-4 * scale
        db = torch.linspace(87, 5, D).repeat(N, 1)
        d_tc = ops.softargmin(out.contiguous(), db.cuda()).cpu()
        d_ref = ops.softargmin(ref.cuda().contiguous(), db.cuda()).cpu()
        assert float(((d_tc - d_ref).abs() / d_ref.abs()).max()) < 1e-4


@pytest.mark.parametrize("fmt", ["tf32", "f16"])
@pytest.mark.parametrize("cfg", [(2, 16, 16, 16, 128, 128, False), (3, 8, 8, 8, 128, 128, True), (1, 4, 16, 16, 64, 128, False)])
def test_conv3d_tc_fused_maxpool(lib, cfg, fmt):
    """MaxPool3d((1,2,2)) fused into the epilogue (relu | 4) == the stand-alone pool of the unfused result, bit for bit,
    for the fp32 output and for both halves of the operand pair."""
    from side_b200 import ops
    N, D, H, W, Cin, Cout, res = cfg
    g = torch.Generator().manual_seed(Cin + H + N)
    dev = torch.device("cuda")
    x = torch.randn(N, Cin, D, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) * 0.05).to(dev)
    sc, sh = (torch.rand(Cout, generator=g) + 0.5).to(dev), (torch.randn(Cout, generator=g) * 0.1).to(dev)
    r = torch.randn(N, D, H, W, Cout, generator=g).to(dev) if res else None
    hi, lo = ops.ncdhw_to_cl_split(x, fmt=fmt)
    wp = ops.conv_tc_prepare(w, fmt=fmt)
    y, _, _ = ops.conv3d_tc(hi, lo, wp, Cout, scale=sc, shift=sh, relu=True, residual=r, full=True, split=False)
    want, whi, wlo = ops.maxpool_hw2_cl(y, full=True, split=True, fmt=fmt)
    got, ghi, glo = ops.conv3d_tc(hi, lo, wp, Cout, scale=sc, shift=sh, relu=True, residual=r, full=True, split=True, pool=True)
    assert tuple(got.shape) == (N, D, H // 2, W // 2, Cout)
    assert torch.equal(got, want) and torch.equal(ghi, whi) and torch.equal(glo, wlo)


# ------------------------------------------------------------------------------------------------
# range of the fp16 operand pairs: fp16 covers 2^-24 .. 65504, the reference's fp32 does not care
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("xs,ws", [(1e-6, 1.0), (1e-3, 1.0), (1e3, 1.0), (1.0, 1e-6), (1.0, 1e-3), (1.0, 1e3), (1.0, 1e5), (1e-3, 1e3)])
def test_conv3d_tc_f16_scaled_operands(lib, xs, ws):
    """Activations / weights scaled by 1e-6 .. 1e5 through the kind::f16 convolution: weights are normalised by an exact power
    of two at preparation, activations down to ~1e-6 keep the 1e-4 bar, and no guard flag fires inside the supported range."""
    from side_b200 import ops
    N, D, H, W, Cin, Cout = 2, 8, 16, 16, 64, 64
    g = torch.Generator().manual_seed(17)
    x = torch.randn(N, Cin, D, H, W, generator=g) * xs
    w = torch.randn(Cout, Cin, 3, 3, 3, generator=g) * (2.0 / (27 * Cout)) ** 0.5 * ws
    ref = _cl(F.conv3d(x.double(), w.double(), padding=1).relu()).numpy()
    dev = torch.device("cuda")
    ops.tc_range_status(dev)
    wp = ops.conv_tc_prepare(w.to(dev), fmt="f16")
    hi, lo = ops.ncdhw_to_cl_split(x.to(dev), fmt="f16")
    y, yh, yl = ops.conv3d_tc(hi, lo, wp, Cout, relu=True, full=True, split=True)
    flags = ops.tc_range_status(dev)
    amax = float(np.abs(ref).max())
    assert rel_err(y.cpu().numpy(), ref) < 1e-4, (xs, ws, flags)       # the fp32 output never depends on the OUTPUT pair's range
    if amax >= 65504.0:
        assert flags & ops.TC_RANGE_SATURATED, (flags, amax)              # ... but the pair handed to the next layer does
    elif 1e-4 < amax < 3e4 and 1e-4 < xs * 4:
        assert flags == 0, (flags, amax)
    if not flags:
        rec = yh.float() + yl.float() / 2048.0
        assert float((rec - y).abs().max()) <= 4e-6 * float(y.abs().max())


def test_f16_range_guard_flags(lib):
    """Saturation (an activation >= 65504) and underflow (a non-zero tensor below 2^-18 everywhere) raise the sticky flags from
    every producer of fp16 pairs; all-zero tensors and ordinary ones do not; reading resets."""
    from side_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 64, 4, 16, 16, generator=g).to(dev)
    ops.tc_range_status(dev)
    ops.ncdhw_to_cl_split(x, fmt="f16")
    ops.ncdhw_to_cl_split(torch.zeros_like(x), fmt="f16")
    assert ops.tc_range_status(dev) == 0
    big = x.clone()
    big[1, 3, 2, 5, 7] = 7.0e4
    ops.ncdhw_to_cl_split(big, fmt="f16")
    assert ops.tc_range_status(dev) == ops.TC_RANGE_SATURATED
    assert ops.tc_range_status(dev) == 0                               # reading cleared it
    ops.ncdhw_to_cl_split(x * 1e-7, fmt="f16")
    assert ops.tc_range_status(dev) == ops.TC_RANGE_UNDERFLOW
    y = torch.randn(2, 4, 16, 16, 64, generator=g).to(dev)
    gate = torch.rand(2, 4, 16, 64, generator=g).to(dev)
    ops.gate_mul_split(y * 1e6, gate, fmt="f16")
    assert ops.tc_range_status(dev) == ops.TC_RANGE_SATURATED
    ops.maxpool_hw2_cl(y * 1e6, full=False, split=True, fmt="f16")
    assert ops.tc_range_status(dev) == ops.TC_RANGE_SATURATED
    ops.gate_mul_split(y, gate, fmt="f16"); ops.maxpool_hw2_cl(y, full=False, split=True, fmt="f16")
    assert ops.tc_range_status(dev) == 0
    # convolution epilogues (voxel-major and role-swapped kernels): the OUTPUT pair saturates
    for Cout in (64, 128):
        w = torch.randn(Cout, 64, 3, 3, 3, generator=g).to(dev) * 0.05
        wp = ops.conv_tc_prepare(w, fmt="f16")
        hi, lo = ops.ncdhw_to_cl_split(x, fmt="f16")
        ops.conv3d_tc(hi, lo, wp, Cout, relu=True, full=False, split=True)
        assert ops.tc_range_status(dev) == 0
        ops.conv3d_tc(hi, lo, wp, Cout, scale=torch.full((Cout,), 1e6, device=dev), relu=True, full=False, split=True)
        assert ops.tc_range_status(dev) == ops.TC_RANGE_SATURATED, Cout
    # the tf32 format has an 8-bit exponent and never reports
    ops.ncdhw_to_cl_split(big, fmt="tf32")
    assert ops.tc_range_status(dev) == 0


@pytest.mark.parametrize("scale", [1e-6, 1e-3, 1e3, 1e5])
def test_aggregate_tc_f16_scaled_volume_with_fallback(lib, scale):
    """Whole aggregation network on a volume scaled by 1e-6 .. 1e5: either the fp16 path meets the 1e-4 bar, or the guard fires
    and the 3xTF32 rerun does (what StereoDetector.process_checked does for the full step)."""
    from side_b200 import ops
    from side_b200.networks.stereo_network import cost_volume
    torch.manual_seed(13)
    m = cost_volume(64).eval()
    for mod in m.modules():
        if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.8, 1.2)
            mod.bias.data.normal_(0, 0.1)
    N, D = 2, 16
    cost = torch.randn(N, 96, D, 16, 16) * scale
    with torch.no_grad():
        ref = m.double().aggregate(cost.double()).float()
    m = m.float().cuda()
    dev = torch.device("cuda")
    rng = float(ref.abs().max())
    with torch.no_grad():
        m.tc_format = "f16"
        ops.tc_range_status(dev)
        out = m.aggregate_tc(cost.cuda())
        flags = ops.tc_range_status(dev)
        if flags:
            m.tc_format = "tf32"
            out = m.aggregate_tc(cost.cuda())
        err = float((out.cpu() - ref).abs().max()) / rng
    assert err < 1e-4, (scale, flags, err)
    if scale >= 1e5:
        assert flags & ops.TC_RANGE_SATURATED            # 1e5 * N(0,1) crosses 65504: the guard must have caught it


def test_stem_block_convolutions_match_direct_stem(lib):
    """DLA stem on the tensor cores (base layer -> 2x2 space-to-depth fp16 pairs, level0 / level1 as 3x3 BLOCK convolutions with
    rearranged weights) against the plain modules in float64: the level-1 output that levels 2-5 start from."""
    from side_b200 import ops
    from side_b200.networks.feature_extraction_dla34 import dla34
    torch.manual_seed(0)
    m = dla34().cuda().eval()
    for seq in (m.base_layer, m.level0, m.level1):
        bn = seq[1]
        with torch.no_grad():
            bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5); bn.weight.uniform_(0.7, 1.3); bn.bias.normal_(0, 0.2)
    x = torch.randn(2, 3, 64, 256, device="cuda")
    old = ops.get_tc_format()
    ops.set_tc_format("f16")
    try:
        with torch.no_grad():
            assert m._stem_tc_ok(x)
            full, hi, lo = m._stem_tc(x)
            md = dla34().double().cuda().eval()
            md.load_state_dict(m.state_dict())
            ref = md.level1(md.level0(md.base_layer(x.double())))                  # [2, 32, 32, 128]
    finally:
        ops.set_tc_format(old)
    got = full[0].permute(0, 3, 1, 2)
    assert got.shape == ref.shape
    assert float((got.double() - ref).abs().max() / ref.abs().max()) < 1e-4
    pair = hi.float() + lo.float() / 2048.0
    assert float((pair[0].permute(0, 3, 1, 2).double() - ref).abs().max() / ref.abs().max()) < 1e-4
    # the block weights are an exact rearrangement: every pixel-domain weight appears once per output sub-position
    w = m.level0[0].weight.detach()
    wb = m._block_weights(w, 1)
    assert wb.shape == (64, 64, 3, 3) and float(wb.abs().sum()) == pytest.approx(4 * float(w.abs().sum()), rel=1e-6)
    wb2 = m._block_weights(m.level1[0].weight.detach(), 2)
    assert wb2.shape == (32, 64, 3, 3) and float(wb2.abs().sum()) == pytest.approx(float(m.level1[0].weight.abs().sum()), rel=1e-6)


def test_cl_concat_matches_torch_cat(lib):
    from side_b200 import ops
    torch.manual_seed(1)
    for dt, widths in ((torch.float16, (64, 64, 128)), (torch.float32, (32, 4, 96)), (torch.float16, (64, 12))):   # last: 24-byte rows -> torch.cat
        ts = [torch.randn(2, 3, 5, 7, w, device="cuda").to(dt) for w in widths]
        assert torch.equal(ops.cl_concat(ts), torch.cat(ts, -1))


@pytest.mark.parametrize("shape", [(5, 16, 4, 4), (3, 32, 4, 4), (2, 48, 4, 4), (4, 7, 3, 5), (1, 2, 1, 1)])
def test_conv3d_c1_cl_one_sample_per_cta_kernel(lib, shape):
    """Conv3d(64 -> 1, 3, padding 1) of the classifier (stereo_network_old.py:161-163): the one-RoI-per-CTA kernel (D*H*W <= 512)
    and the generic gather kernel (larger volumes) against float64, <= 1e-5 of the output range."""
    from side_b200 import ops
    N, D, H, W = shape
    g = torch.Generator().manual_seed(N * 1000 + D)
    y = torch.randn(N, D, H, W, 64, generator=g)
    w = torch.randn(1, 64, 3, 3, 3, generator=g) * 0.1
    out = ops.conv3d_c1_cl(y.cuda(), w.cuda())
    ref = F.conv3d(y.permute(0, 4, 1, 2, 3).double(), w.double(), padding=1)[:, 0]
    assert tuple(out.shape) == (N, D, H, W)
    assert rel_err(out.cpu().numpy(), ref.numpy()) < 1e-5
