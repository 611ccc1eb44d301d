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
