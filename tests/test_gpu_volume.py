"""GPU parity: full-image concat / group-correlation volumes (north_star item 1).
PARITY UNPINNED by the reference (no call site): the checker is the oracle restatement."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

from oracle import c_oracle as co  # noqa: E402
from oracle import torch_port as tp  # noqa: E402


@pytest.mark.parametrize("cfg", [(2, 8, 5, 24, 6), (1, 3, 9, 33, 4), (1, 4, 20, 320, 48), (2, 2, 8, 12, 12), (1, 5, 3, 7, 1)])
def test_concat_bit_exact_and_backward(lib, cfg):
    from side_b200 import ops
    B, C, H, W, D = cfg
    rng = np.random.default_rng(B + C + W)
    L = rng.standard_normal((B, C, H, W)).astype(np.float32)
    R = rng.standard_normal((B, C, H, W)).astype(np.float32)
    a, b = torch.from_numpy(L).cuda().requires_grad_(True), torch.from_numpy(R).cuda().requires_grad_(True)
    v = ops.concat_volume(a, b, D)
    assert np.array_equal(v.detach().cpu().numpy(), co.concat_volume(L, R, D))
    assert torch.equal(v[:, :, 0], torch.cat((a, b), 1))          # D=1 slice == stereo-head concat (:348)
    g = torch.from_numpy(rng.standard_normal(tuple(v.shape)).astype(np.float32))
    ga, gb = torch.autograd.grad(v, (a, b), g.cuda())
    la, lb = torch.from_numpy(L).requires_grad_(True), torch.from_numpy(R).requires_grad_(True)
    ra, rb = torch.autograd.grad(tp.concat_volume(la, lb, D), (la, lb), g)
    assert rel_err(ga.cpu().numpy(), ra.numpy()) < 1e-5 and rel_err(gb.cpu().numpy(), rb.numpy()) < 1e-5


@pytest.mark.parametrize("cfg", [(2, 8, 5, 24, 6, 4), (1, 6, 9, 33, 4, 2), (1, 64, 6, 320, 48, 8), (1, 32, 4, 40, 5, 2), (1, 12, 3, 16, 3, 1)])
def test_gwc_forward_backward(lib, cfg):
    from side_b200 import ops
    B, C, H, W, D, G = cfg
    rng = np.random.default_rng(C + W + G)
    L = rng.standard_normal((B, C, H, W)).astype(np.float32)
    R = rng.standard_normal((B, C, H, W)).astype(np.float32)
    a, b = torch.from_numpy(L).cuda().requires_grad_(True), torch.from_numpy(R).cuda().requires_grad_(True)
    v = ops.gwc_volume(a, b, D, G)
    assert rel_err(v.detach().cpu().numpy(), co.gwc_volume(L, R, D, G)) < 1e-4     # 1e-4 relative for fp32 correlation
    g = torch.from_numpy(rng.standard_normal(tuple(v.shape)).astype(np.float32))
    ga, gb = torch.autograd.grad(v, (a, b), g.cuda())
    la, lb = torch.from_numpy(L).requires_grad_(True), torch.from_numpy(R).requires_grad_(True)
    ra, rb = torch.autograd.grad(tp.gwc_volume(la, lb, D, G), (la, lb), g)
    assert rel_err(ga.cpu().numpy(), ra.numpy()) < 1e-4 and rel_err(gb.cpu().numpy(), rb.numpy()) < 1e-4


def test_full_size_properties(lib):
    """C=64, D=48, 96x320 (755 MB / pair): shift-gather identities hold on the whole volume."""
    from side_b200 import ops
    torch.manual_seed(0)
    L, R = torch.randn(1, 64, 96, 320, device="cuda"), torch.randn(1, 64, 96, 320, device="cuda")
    v = ops.concat_volume(L, R, 48)
    for d in (0, 1, 7, 47):
        assert torch.equal(v[:, :64, d, :, d:], L[..., d:]) and torch.equal(v[:, 64:, d, :, d:], R[..., :320 - d])
        assert d == 0 or (v[:, :, d, :, :d] == 0).all()
    w = ops.gwc_volume(L, R, 48, 8)
    ref = (L[..., 5:] * R[..., :315]).view(1, 8, 8, 96, 315).mean(2)
    assert (w[:, :, 5, :, 5:] - ref).abs().max().item() < 1e-5
    assert (w[:, :, 5, :, :5] == 0).all()
