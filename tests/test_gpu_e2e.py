"""GPU parity, end to end: the drop-in stereo_network on the CUDA kernels vs (a) golden outputs produced by the
reference's own stereo_network_old with the same state dict, (b) the CPU port executed live on this box."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


def _state_sha(model):
    return hashlib.sha256(np.concatenate([v.detach().cpu().numpy().reshape(-1).astype(np.float64)
                                          for v in model.state_dict().values()]).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def model():
    from side_b200.networks import get_pose_net
    from side_b200.utils.synthetic import HEADS, realistic_init
    torch.manual_seed(0)
    return realistic_init(get_pose_net(34, HEADS, 256), seed=1).eval()


@pytest.fixture(autouse=True)
def _fp32_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_e2e_against_reference_golden(lib, model):
    from side_b200.decode import ddd_decode
    from side_b200.utils.synthetic import make_batch, make_boxes
    g = golden("e2e_small")
    if _state_sha(model) != str(g["state_sha256"]):
        pytest.skip("seeded init differs from the build container's (different torch build): golden not applicable")
    batch = make_batch(1, 64, 1280, seed=3)
    assert hashlib.sha256(batch['input'].numpy().tobytes()).hexdigest() == str(g["input_sha256"])
    m = model.cuda()
    cb = {k: v.cuda() for k, v in batch.items()}
    left, right, shape = make_boxes(1, 6, seed=9, W4=320, H4=16)
    with torch.no_grad():
        z = m(cb, True, None, 1.0)[0]
        zt = m(cb, True, (left.cuda(), right.cuda(), shape), 1.0)[0]
    pos = torch.from_numpy(g["pos"]).cuda()
    for k in ("hm", "wh", "reg"):
        ref = g[k]
        err = np.abs(z[k].cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < 1e-3, (k, err)                     # ~120 cuDNN layers between input and heads
    for k in ("kept_type", "dim", "orien"):
        mine = z[k].reshape(z[k].shape[0], z[k].shape[1], -1)[:, :, pos].cpu().numpy()
        assert np.abs(mine - g[k + "_s"]).max() / np.abs(g[k + "_s"]).max() < 1e-3, k
    dt = zt['depth'].cpu().numpy()
    assert np.abs(dt - g["depth_target"]).max() / np.abs(g["depth_target"]).max() < 2e-3
    d = z['depth'].cpu().numpy()
    close = np.abs(d - g["depth"]) <= 2e-3 * np.abs(g["depth"]).max()
    assert close.mean() > 0.9, "inference depth: %.0f%% of rows agree" % (100 * close.mean())
    hm = z['hm'].clone().sigmoid_()
    det, detr, info = ddd_decode(hm, z['kept_type'], z['dim'], z['orien'], wh=z['wh'], reg=z['reg'], grid_size=28, K=100)
    same = np.isclose(det.cpu().numpy()[0, :, :2], g["det"][0, :, :2], atol=1e-2).all(1)
    assert same.mean() > 0.9, "decoded centres: %.0f%% of rows agree" % (100 * same.mean())


def test_e2e_against_live_cpu_port(lib, model):
    """Same weights, same inputs: product on the GPU vs the reference-style port on this box's CPU."""
    from oracle import torch_port
    from side_b200.utils.synthetic import make_batch, make_boxes
    batch = make_batch(2, 64, 1280, seed=11)
    left, right, shape = make_boxes(2, 5, seed=3, W4=320, H4=16)
    cpu_model = model.cpu()
    with torch.no_grad(), torch_port.reference_ops():
        zr = cpu_model(batch, True, (left, right, shape), 1.0)[0]
    m = model.cuda()
    with torch.no_grad():
        z = m({k: v.cuda() for k, v in batch.items()}, True, (left.cuda(), right.cuda(), shape), 1.0)[0]
    for k in zr:
        ref = zr[k].numpy()
        err = np.abs(z[k].cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < 2e-3, (k, err)


def test_training_step_gradients_flow(lib, model):
    """Config #5 shape of work at small size: forward with GT RoIs, L1 depth loss + head losses, backward through
    DCN / instance volume / soft-argmin kernels; gradients finite and non-zero on the hot-path parameters."""
    from side_b200.utils.synthetic import make_batch, make_boxes
    m = model.cuda().train()
    batch = {k: v.cuda() for k, v in make_batch(2, 64, 1280, seed=5).items()}
    left, right, shape = make_boxes(2, 4, seed=1, W4=320, H4=16)
    try:
        z = m(batch, True, (left.cuda(), right.cuda(), shape), 1.0)[0]
        loss = z['depth'].abs().mean() + sum(z[k].pow(2).mean() for k in ("hm", "wh", "reg"))
        loss.backward()
        names = ["feature_extraction.ida_up.node_2.conv.weight", "feature_extraction.ida_up.node_2.conv.conv_offset_mask.weight",
                 "feaRuduce.0.weight", "depth_estimator.classify.3.weight", "feature_extraction.base.level2.tree1.conv1.weight"]
        params = dict(m.named_parameters())
        for n in names:
            g = params[n].grad
            assert g is not None and torch.isfinite(g).all() and g.abs().sum().item() > 0, n
    finally:
        m.zero_grad(set_to_none=True)
        m.eval()


def test_cuda_graph_capture_matches_eager(lib, model):
    from side_b200.engine import StereoDetector
    from side_b200.utils.synthetic import make_batch
    m = model.cuda().eval()
    det = StereoDetector(m)
    b1 = {k: v.cuda() for k, v in make_batch(1, 64, 1280, seed=21).items()}
    b2 = {k: v.cuda() for k, v in make_batch(1, 64, 1280, seed=22).items()}
    eager = [t.clone() for t in det.process(b2)]
    det.capture(b1)
    out = det.replay(b2)
    torch.cuda.synchronize()
    for a, b in zip(eager, out):
        assert torch.allclose(a, b, atol=1e-4, rtol=1e-4)
    assert det.launches_per_step > 30


def test_bench_configuration_against_cpu_port(lib, model):
    """the EXACT configuration bench.py times -- 384x1280, micro-batch 8, K = 100, DCN on tcgen05 3xTF32,
    every tensor-core convolution on 3xFP16 pairs, separable volume, tensor-core DLA / heads / stem -- against the
    reference-style port on this box's CPU with the same weights and inputs (2 of the 8 pairs: ~10 s of CPU)."""
    from oracle import parity, torch_port
    from side_b200 import ops
    from side_b200.utils.synthetic import make_batch
    n_ref = 2
    batch = make_batch(8, 384, 1280, seed=1234)
    cpu_model = model.cpu().eval()
    with torch.no_grad(), torch_port.reference_ops():
        zr = cpu_model({k: v[:n_ref] for k, v in batch.items()}, True, None, 1.0)[0]
    m = model.cuda().eval()
    old_fmt, old_prec = ops.get_tc_format(), ops.get_dcn_precision()
    ops.set_tc_format("f16"); ops.set_dcn_precision("3xtf32")
    try:
        dev = torch.device("cuda")
        ops.tc_range_status(dev)
        with torch.no_grad():
            z = m({k: v.cuda() for k, v in batch.items()}, True, None, 1.0)[0]
        assert ops.tc_range_status(dev) == 0, "fp16 range guard fired on the bench configuration"
    finally:
        ops.set_tc_format(old_fmt); ops.set_dcn_precision(old_prec)
    rep = parity.compare({k: v[:n_ref] for k, v in z.items()}, zr, K=100)
    print("bench-configuration parity:", rep)
    assert rep["heads_max_err"] < 1e-3, rep                      # per head, relative to the head's range (~130 layers deep)
    assert rep["topk_agreement"] >= 0.98, rep                    # (class, peak index) keys found by both runs
    assert rep["depth_median_rel"] < 1e-4, rep
    assert rep["depth_frac_within_1e3"] >= 0.98, rep             # per matched detection |d - d_ref| / d_ref <= 1e-3
