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


def test_training_glue_vs_reference_losses_and_gt_rois(R):
    """side_b200.training against the reference's own loss classes (models/losses.py:114-198) and a literal restatement of
    ModelWithLoss.forward's RoI construction (modules/stereoTrainer.py:41-63; that module itself needs `progress`, absent here)."""
    import importlib
    from side_b200 import training as T
    from side_b200.utils.synthetic import make_targets
    losses = importlib.import_module("models.losses")
    t = make_targets(2, 8, seed=3)
    torch.manual_seed(1)
    out = {'hm': torch.randn(2, 3, 96, 320) - 2, 'wh': torch.rand(2, 3, 96, 320) * 20, 'reg': torch.rand(2, 3, 96, 320),
           'dim': torch.randn(2, 3, 96, 320), 'orien': torch.randn(2, 2, 96, 320), 'kept_type': torch.randn(2, 168, 96, 320)}
    hm = torch.clamp(torch.sigmoid(out['hm']), 1e-4, 1 - 1e-4)
    assert torch.allclose(T.focal_loss(hm, t['hm']), losses.FocalLoss()(hm, t['hm']), rtol=1e-6)
    assert torch.allclose(T.focal_loss(hm, torch.zeros_like(hm)), losses.FocalLoss()(hm, torch.zeros_like(hm)), rtol=1e-6)
    assert torch.allclose(T.reg_l1(out['wh'], t['rot_mask'], t['ind'], t['wh']), losses.L1Loss()(out['wh'], t['rot_mask'], t['ind'], t['wh']), rtol=1e-6)
    tgt = T.kept_label(t['kept'], t['wh'], 28)
    assert torch.allclose(T.cross_loss(out['kept_type'][:, :112], t['ind'], tgt[:, :, 0]),
                          losses.CrossLoss()(out['kept_type'][:, :112], t['rot_mask'], t['ind'], tgt[:, :, 0].unsqueeze(2)), rtol=1e-6)
    # stereoTrainer.py:41-63, restated literally
    xs, ys = t['ind_float'] % 320, t['ind_float'] // 320
    wh, reg = t['wh'], t['reg']
    xs_right = xs + reg[:, :, 1]
    xs, ys = xs + reg[:, :, 0], ys + reg[:, :, 2]
    center = torch.cat([xs.unsqueeze(2), ys.unsqueeze(2)], dim=2)
    center_right = torch.cat([xs_right.unsqueeze(2), ys.unsqueeze(2)], dim=2)
    bidx = torch.tensor([n for n in range(xs.shape[0])], dtype=torch.float32).unsqueeze(1).unsqueeze(2).repeat(1, xs.shape[1], 1)
    bbox = torch.zeros((xs.shape[0], xs.shape[1], 5)); bbox_right = torch.zeros((xs.shape[0], xs.shape[1], 5))
    bbox[:, :, 1:3] = center - 0.5 * wh[:, :, [0, 2]]; bbox[:, :, 3:5] = center + 0.5 * wh[:, :, [0, 2]]
    bbox_right[:, :, 1:3] = center_right - 0.5 * wh[:, :, [1, 2]]; bbox_right[:, :, 3:5] = center_right + 0.5 * wh[:, :, [1, 2]]
    bbox[:, :, 0:1], bbox_right[:, :, 0:1] = bidx, bidx
    keep = torch.sum(bbox.view(-1, 5)[:, 1:5], dim=1) > 0
    bl, br, shp, k8 = T.gt_rois(t, 320)
    assert torch.equal(k8.bool(), keep) and tuple(shp) == tuple(bbox.shape)
    assert torch.equal(bl[keep], bbox.view(-1, 5)[keep]) and torch.equal(br[keep], bbox_right.view(-1, 5)[keep])


def test_voxel_variant_reproduces_reference_network(R):
    """stereo_network_new (voxel / PointNet head): same state-dict keys, and the product's module tree with the port ops ==
    the reference network on the same weights (quarter-size configuration, ground-truth RoIs)."""
    import importlib
    import warnings
    from oracle import torch_port
    from oracle.gen_golden import voxel_new_case
    from side_b200.networks import stereo_network_new as sn
    from side_b200.utils.synthetic import HEADS
    rn = importlib.import_module("models.networks.stereo_network_new")
    c = voxel_new_case()
    old = (rn.input_h, rn.input_w, sn.input_h, sn.input_w)
    rn.input_h, rn.input_w = float(c["H_in"]), float(c["W_in"])
    sn.input_h, sn.input_w = c["H_in"], c["W_in"]
    try:
        torch.manual_seed(3)
        ref = rn.get_pose_net(34, HEADS, 256).eval()
        mine = sn.get_pose_net(34, HEADS, 256).eval()
        assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
        mine.load_state_dict(ref.state_dict())
        t = lambda k: torch.from_numpy(c[k])
        g = torch.Generator().manual_seed(5)
        batch = {'input': torch.randn(2, 3, c["H_in"], c["W_in"], generator=g), 'input_right': torch.randn(2, 3, c["H_in"], c["W_in"], generator=g),
                 'fb': t("fb"), 'p2': t("p2"), 'p3': t("p3"), 'trans': t("trans"), 'trans_inv': t("trans_inv")}
        target = (t("left"), t("right"), torch.Size([2, 50, 1]))
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            zr = ref(batch, True, target)[0]
            with torch_port.reference_ops():
                zm = mine(batch, True, target)[0]
        for k in zr:
            err = (zr[k] - zm[k]).abs().max().item() / max(zr[k].abs().max().item(), 1e-30)
            assert err < 1e-4, (k, err)
    finally:
        rn.input_h, rn.input_w, sn.input_h, sn.input_w = old
