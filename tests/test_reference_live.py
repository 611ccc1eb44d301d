"""CPU, build container only (skipped where /root/reference is absent): the oracle and the product's module tree
against the UNMODIFIED reference executed live."""
import numpy as np
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def R():
    return ref_loader.load()


def test_state_dict_keys_match_reference(R):
    from side_b200.networks import get_pose_net
    from side_b200.utils.synthetic import HEADS
    mine = get_pose_net(34, HEADS, 256).state_dict()
    ref = R.net.get_pose_net(34, HEADS, 256).state_dict()
    assert list(mine.keys()) == list(ref.keys())
    assert all(mine[k].shape == ref[k].shape for k in mine)


def test_port_reproduces_reference_network(R):
    """The product's nn.Module tree driven by the reference-style CPU ops == stereo_network_old, same state dict."""
    from oracle import torch_port
    from side_b200.networks import get_pose_net
    from side_b200.utils.synthetic import HEADS, make_batch, make_boxes, realistic_init
    torch.manual_seed(0)
    mine = realistic_init(get_pose_net(34, HEADS, 256), seed=1).eval()
    ref = R.net.get_pose_net(34, HEADS, 256).eval()
    ref.load_state_dict(mine.state_dict())
    batch = make_batch(1, 64, 1280, seed=3)
    left, right, shape = make_boxes(1, 6, seed=9, W4=320, H4=16)
    with torch.no_grad():
        zr = ref(batch, True, (left, right, shape), 1.0)[0]
        with torch_port.reference_ops():
            zm = mine(batch, True, (left, right, shape), 1.0)[0]
    for k in zr:
        err = (zr[k] - zm[k]).abs().max().item() / zr[k].abs().max().item()
        assert err < 1e-4, (k, err)


def test_decode_port_vs_reference(R):
    from oracle import torch_port as tp
    torch.manual_seed(4)
    hm = torch.randn(2, 3, 24, 40) * 1.5 - 2.19
    wh = torch.rand(2, 3, 24, 40) * 30
    reg = torch.rand(2, 3, 24, 40)
    bk, brk, shape = R.decode.bbox_decode(hm, wh, reg, K=20)
    o = tp.bbox_decode_raw(hm, wh, reg, K=20)
    keep = o["keep"].bool()
    assert torch.equal(o["bbox"].view(-1, 5)[keep], bk) and torch.equal(o["bbox_right"].view(-1, 5)[keep], brk)


def test_c_oracle_roi_align_bit_exact_vs_torchvision():
    import torchvision.ops as tvo
    from oracle import c_oracle as co
    torch.manual_seed(0)
    feat = torch.randn(2, 8, 24, 40)
    N = 64
    x1 = torch.rand(N) * 44 - 4; y1 = torch.rand(N) * 26 - 3
    rois = torch.stack([torch.randint(0, 2, (N,)).float(), x1, y1, x1 + torch.rand(N) * 25, y1 + torch.rand(N) * 14], 1)
    for P in (16, 7):
        ref = tvo.roi_align(feat, rois, (P, P), 1.0, 2).numpy()
        assert np.array_equal(ref, co.roi_align(feat.numpy(), rois.numpy(), P))
