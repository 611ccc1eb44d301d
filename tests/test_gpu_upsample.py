"""GPU parity: depth-wise ConvTranspose2d of IDAUp (feature_extraction_dla34.py:370-373) vs the ATen op the
reference calls (F.conv_transpose2d, groups = channels), forward and backward."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [(2, 64, 24, 80, 2), (1, 64, 12, 40, 4), (3, 5, 7, 9, 2), (1, 3, 5, 6, 8), (2, 4, 6, 5, 1), (1, 2, 4, 4, 3)])
def test_dw_deconv_forward_backward(lib, cfg):
    from side_b200 import ops
    B, C, H, W, f = cfg
    torch.manual_seed(C + f)
    x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
    w = torch.randn(C, 1, 2 * f, 2 * f, device="cuda", requires_grad=True)
    y = ops.dw_deconv(x, w, f, f // 2)
    xr, wr = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    ref = F.conv_transpose2d(xr, wr, None, stride=f, padding=f // 2, groups=C)
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    g = torch.randn_like(ref)
    gx, gw = torch.autograd.grad(y, (x, w), g)
    rx, rw = torch.autograd.grad(ref, (xr, wr), g)
    assert (gx - rx).abs().max().item() <= 1e-4 * rx.abs().max().item()
    assert (gw - rw).abs().max().item() <= 1e-4 * rw.abs().max().item()


def test_bilinear_init_module_matches_reference_layer(lib):
    """DepthwiseUp keeps ConvTranspose2d's parameters / state-dict key and the bilinear fill (reference :333-342)."""
    from side_b200.networks.feature_extraction_dla34 import DepthwiseUp, fill_up_weights
    up = DepthwiseUp(8, 8, 4, stride=2, padding=1, output_padding=0, groups=8, bias=False)
    fill_up_weights(up)
    ref = torch.nn.ConvTranspose2d(8, 8, 4, stride=2, padding=1, output_padding=0, groups=8, bias=False)
    ref.load_state_dict(up.state_dict())
    x = torch.randn(2, 8, 9, 11)
    assert torch.allclose(up.cuda()(x.cuda()).cpu(), ref(x), atol=1e-6)


@pytest.mark.parametrize("cfg", [(2, 64, 24, 80, 2), (3, 128, 12, 40, 4), (1, 64, 12, 40, 8), (2, 96, 5, 7, 2)])
def test_idaup_fused_step_is_bit_identical_to_the_three_kernel_sequence(lib, cfg):
    """ops.idaup_fuse_cl (up_k + skip addition + channels-last + fp16 operand split, feature_extraction_dla34.py:380-386) against
    dw_deconv -> torch add -> ncdhw_to_cl_split: the same bits in all three outputs."""
    from side_b200 import ops
    B, C, H, W, f = cfg
    g = torch.Generator().manual_seed(B * 100 + C + f)
    x = torch.randn(B, C, H, W, generator=g).cuda()
    skip = torch.randn(B, C, H * f, W * f, generator=g).cuda()
    w = torch.randn(C, 1, 2 * f, 2 * f, generator=g).cuda()
    ops.set_tc_format("f16")
    full0, hi0, lo0 = ops.ncdhw_to_cl_split((ops.dw_deconv(x, w, f, f // 2) + skip).unsqueeze(2), want_full=True)
    full1, hi1, lo1 = ops.idaup_fuse_cl(x, w, skip, f)
    assert full1.shape == full0.shape and hi1.dtype == torch.float16
    assert torch.equal(full0, full1) and torch.equal(hi0, hi1) and torch.equal(lo0, lo1)
    ref = torch.nn.functional.conv_transpose2d(x.double(), w.double(), stride=f, padding=f // 2, groups=C) + skip.double()
    got = full1[:, 0, :, :, :C].permute(0, 3, 1, 2).double()
    assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_idaup_forward_fused_equals_unfused(lib):
    """IDAUp.forward with the fused step == the module sequence (proj DCN -> up -> add -> node DCN), same bits."""
    from side_b200 import ops
    from side_b200.networks.feature_extraction_dla34 import IDAUp
    torch.manual_seed(3)
    ops.set_tc_format("f16")
    ops.set_dcn_precision("3xfp16")
    m = IDAUp(64, [64, 128, 256], [1, 2, 4]).cuda().eval()
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
            if hasattr(mod, "conv_offset_mask"):
                mod.conv_offset_mask.weight.normal_(0, 0.01); mod.conv_offset_mask.bias.normal_(0, 0.1)
        layers = [torch.randn(4, 64, 48, 160).cuda(), torch.randn(4, 128, 24, 80).cuda(), torch.randn(4, 256, 12, 40).cuda()]
        a = [t.clone() for t in layers]
        b = [t.clone() for t in layers]
        m.fuse_up_add = True
        m(a, 0, 3)
        m.fuse_up_add = False
        m(b, 0, 3)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
