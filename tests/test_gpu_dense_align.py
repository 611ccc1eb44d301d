"""GPU parity: dense photometric alignment (csrc/dense_align.cu, side_b200/dense_align.py) against the reference's golden
vectors (tests/golden/dense_align.npz, produced by executing dense_align.py itself) and against the C oracle."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err

pytestmark = pytest.mark.gpu

from oracle import c_oracle as co  # noqa: E402
from oracle.gen_golden import dense_align_case  # noqa: E402  (pure numpy: seeded inputs, no reference access)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_prepare_image_vs_golden(lib):
    from side_b200 import dense_align as da
    g = golden("dense_align")
    img_l, img_r, _, opt, _, _, _ = dense_align_case()
    for img, key in ((img_l, "im_l_s"), (img_r, "im_r_s")):
        p = da.prepare_image(img, opt.mean, opt.std, "cuda")
        planar = p.planar()[0].cpu().numpy()
        assert planar.shape == (3, 2 * img.shape[0], 2 * img.shape[1])
        assert np.abs(planar.reshape(3, -1)[:, g["im_pos"]] - g[key]).max() < 2e-6
        assert np.array_equal(planar, co.da_prep_u8(img, opt.mean, opt.std))        # bit-exact vs the C restatement
        assert float(p.data[..., 3].abs().max()) == 0.0


def test_sample_bit_exact(lib):
    from side_b200 import dense_align as da
    g = golden("dense_align")
    img_l, _, calib, _, box, borders, poses = dense_align_case()
    H, W = img_l.shape[:2]
    uvz, wgt = da.sample(calib, 2, 2 * H, 2 * W, dev(box * 2), dev(poses), dev(borders * 2))
    assert np.array_equal(wgt.cpu().numpy(), g["weight"])
    assert np.array_equal(uvz.cpu().numpy(), g["uvz"])


def test_sample_random_boxes_vs_oracle(lib):
    """Python-slice edge cases: boxes hanging over the image, negative starts (wrap like Python slices), empty ranges, coarse
    steps (extent > 112), boxes seen from every side."""
    from side_b200 import dense_align as da
    import types
    rng = np.random.RandomState(3)
    n = 64
    x1 = rng.uniform(-30, 600, n); w = rng.uniform(2, 260, n); y1 = rng.uniform(-20, 150, n); h = rng.uniform(2, 150, n)
    box = np.stack([x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
    borders = np.stack([x1 + rng.uniform(0, 5, n), x1 + w - rng.uniform(0, 5, n)], 1).astype(np.float32)
    poses = np.stack([rng.uniform(-12, 12, n), rng.uniform(1, 2, n), rng.uniform(4, 60, n), rng.uniform(1.4, 2, n),
                      rng.uniform(1.3, 2, n), rng.uniform(3, 5, n), rng.uniform(-3.2, 3.2, n)], 1).astype(np.float32)
    p2 = np.array([[360.77, 0, 304.78, 22.43], [0, 360.77, 86.4, 0.1], [0, 0, 1, 0.00275]], np.float32)
    calib = types.SimpleNamespace(p2=p2)
    f_h, f_w = 192, 640
    uvz, wgt = da.sample(calib, 1, f_h, f_w, dev(box), dev(poses), dev(borders))
    ou, ow, cnt = co.da_sample(box, borders, poses, float(p2[0, 0]), float(p2[0, 2]), float(p2[1, 2]), f_h, f_w)
    assert cnt.max() > 0 and (cnt == 0).any()
    assert np.array_equal(wgt.cpu().numpy(), ow)
    assert np.array_equal(uvz.cpu().numpy(), ou)


@pytest.mark.parametrize("align", [False, True])
def test_enumeration_depth(lib, align):
    from side_b200 import dense_align as da
    g = golden("dense_align")
    img_l, img_r, _, opt, _, _, _ = dense_align_case()
    L, R = da.prepare_image(img_l, opt.mean, opt.std, "cuda"), da.prepare_image(img_r, opt.mean, opt.std, "cuda")
    da.ALIGN_CORNERS = align
    try:
        best, err, idx = da.enumeration_depth(L.planar(), R.planar(), dev(g["uvz"]), dev(g["weight"]), dev(g["depth_enum"]),
                                              float(g["fb"]), return_error=True)
    finally:
        da.ALIGN_CORNERS = False
    if not align:                       # what the reference computes under this image's torch: the golden vectors
        assert rel_err(err.cpu().numpy(), g["err_sum"]) < 1e-4          # north_star: 1e-4 relative for fp32 reductions
        assert np.array_equal(best.cpu().numpy(), g["best_depth"])
        assert np.array_equal(idx.cpu().numpy(), g["err_sum"].argmin(0))
    oe, ob, oi = co.da_enum(L.planar()[0].cpu().numpy(), R.planar()[0].cpu().numpy(), g["uvz"], g["weight"], g["depth_enum"],
                            float(g["fb"]), align_corners=align)
    assert rel_err(err.cpu().numpy(), oe) < 1e-5
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(best.cpu().numpy(), ob)
    # and torch's own grid_sample on the GPU in the reference's formulation, for the convention that has no golden vector
    uvz, wgt, de = dev(g["uvz"]), dev(g["weight"]), dev(g["depth_enum"])
    Lp, Rp = L.planar(), R.planar()
    f_h, f_w = float(Lp.shape[2]) - 1, float(Lp.shape[3]) - 1
    fb = float(g["fb"])
    ref = torch.empty_like(err)
    for i in range(de.shape[0]):
        dis = de[i].reciprocal() * fb
        dd = (uvz[:, :, 2] / fb + dis.reciprocal().unsqueeze(1)).reciprocal()
        gl = torch.stack([(uvz[:, :, 0] - f_w / 2) / (f_w / 2), (uvz[:, :, 1] - f_h / 2) / (f_h / 2)], -1).unsqueeze(0)
        gr = torch.stack([(uvz[:, :, 0] - dd - f_w / 2) / (f_w / 2), (uvz[:, :, 1] - f_h / 2) / (f_h / 2)], -1).unsqueeze(0)
        e = torch.nn.functional.grid_sample(Lp, gl, padding_mode='border', align_corners=align) - \
            torch.nn.functional.grid_sample(Rp, gr, padding_mode='border', align_corners=align)
        ref[i] = (e[0] * wgt.unsqueeze(0)).abs().sum((0, 2))
    assert rel_err(err.cpu().numpy(), ref.cpu().numpy()) < 1e-4


def test_align_parallel_vs_golden(lib):
    from side_b200 import dense_align as da
    g = golden("dense_align")
    img_l, img_r, calib, opt, box, borders, poses = dense_align_case()
    status, best_dis = da.align_parallel(calib, opt, img_l, img_r, dev(box), dev(borders), dev(poses))
    assert np.array_equal(status.cpu().numpy(), g["status"])
    assert np.abs(best_dis.cpu().numpy() - g["best_dis"]).max() < 1e-4 * np.abs(g["best_dis"]).max()
    # no RoI with a valid pixel: the reference returns zeros and the initial disparity (dense_align.py:281-283)
    status, dis = da.align_parallel(calib, opt, img_l, img_r, dev(box[1:2]), dev(borders[1:2]), dev(poses[1:2]))
    f = calib.p2[0, 0] * 2
    bl = (calib.p2[0, 3] - calib.p3[0, 3]) * 2 / f
    assert status.cpu().numpy().tolist() == [0.0]
    assert np.allclose(dis.cpu().numpy(), f * bl / poses[1:2, 2], rtol=1e-6)


def test_full_size_recovers_known_shift(lib):
    """BASELINE size (384 x 1280 raw -> 768 x 2560 sampled): the right image is the left one shifted by an integer disparity,
    so the photometric minimum must sit at the depth whose disparity is that shift (a size-independent property), and the
    kernel must agree with the oracle on a subset of RoIs."""
    from side_b200 import dense_align as da
    import types
    rng = np.random.RandomState(11)
    H, W, shift = 384, 1280, 12
    small = rng.rand(H // 8 + 1, W // 8 + 1, 3).astype(np.float32)
    img = np.kron(small, np.ones((8, 8, 1), np.float32))[:H, :W]
    k = np.ones(9, np.float32) / 9
    for ax in (0, 1):
        img = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), ax, img)
    img_l = (np.clip(img, 0, 1) * 255).astype(np.uint8)
    img_r = np.ascontiguousarray(np.roll(img_l, -shift, axis=1))
    p2 = np.array([[721.54, 0, 609.56, 44.86], [0, 721.54, 172.85, 0.216], [0, 0, 1, 0.00275]], np.float32)
    p3 = p2.copy(); p3[0, 3] = -339.52
    calib = types.SimpleNamespace(p2=p2, p3=p3)
    opt = types.SimpleNamespace(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    fb = 721.54 * 0.5327
    z = fb / shift                                             # depth whose disparity is `shift` raw pixels (~32 m)
    n = 48
    xs = rng.uniform(-8, 8, n)
    poses = np.stack([xs, np.full(n, 1.6), np.full(n, z) + rng.uniform(-3, 3, n), np.full(n, 1.6), np.full(n, 1.5),
                      np.full(n, 3.9), rng.uniform(-1.5, 1.5, n)], 1).astype(np.float32)
    u = 721.54 * xs / poses[:, 2] + 609.56
    v = 721.54 * 1.6 / poses[:, 2] + 172.85
    box = np.stack([u - 60, v - 50, u + 60, v + 4], 1).astype(np.float32)
    borders = np.stack([u - 55, u + 55], 1).astype(np.float32)
    status, best_dis = da.align_parallel(calib, opt, img_l, img_r, dev(box), dev(borders), dev(poses))
    status, best_dis = status.cpu().numpy(), best_dis.cpu().numpy()
    ok = status == 1
    assert ok.sum() >= n // 2
    # best_dis = disparity at the object centre + 0.5; the visible surface is up to half a car length in front of the centre
    assert np.abs(best_dis[ok] - 0.5 - shift).max() < 1.5
    # oracle on 3 RoIs at full size
    L, R = da.prepare_image(img_l, opt.mean, opt.std, "cuda"), da.prepare_image(img_r, opt.mean, opt.std, "cuda")
    sel = np.nonzero(ok)[0][:3]
    f, cx, cy = p2[0, 0] * 2, p2[0, 2] * 2, p2[1, 2] * 2
    ou, ow, _ = co.da_sample(box[sel] * 2, borders[sel] * 2, poses[sel], float(f), float(cx), float(cy), 2 * H, 2 * W)
    de = (np.linspace(-5, 5, 20, dtype=np.float32)[:, None] + poses[sel, 2][None]).astype(np.float32)
    bl = (p2[0, 3] - p3[0, 3]) * 2 / f
    best, err, idx = da.enumeration_depth(L, R, dev(ou), dev(ow), dev(de), float(f * bl), return_error=True)
    oe, ob, oi = co.da_enum(L.planar()[0].cpu().numpy(), R.planar()[0].cpu().numpy(), ou, ow, de, float(f * bl))
    assert rel_err(err.cpu().numpy(), oe) < 1e-5
    assert np.array_equal(idx.cpu().numpy(), oi)


def test_rejects_cpu_tensors(lib):
    from side_b200 import dense_align as da
    with pytest.raises(RuntimeError):
        da.enumeration_depth(torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8), torch.zeros(1, 4, 3), torch.zeros(1, 4),
                             torch.ones(2, 1), 1.0)


def test_no_rois_and_no_pixels(lib):
    """Empty RoI set and RoIs without a single valid pixel: shapes and the reference's early-exit values."""
    from side_b200 import dense_align as da
    img_l, img_r, calib, opt, box, borders, poses = dense_align_case()
    e = lambda n: torch.zeros((0, n), device="cuda")
    status, dis = da.align_parallel(calib, opt, img_l, img_r, e(4), e(2), e(7))
    assert status.shape == (0,) and dis.shape == (0,)
    L = da.prepare_image(img_l, opt.mean, opt.std, "cuda")
    best, err, idx = da.enumeration_depth(L, L, torch.zeros(2, 0, 3, device="cuda"), torch.zeros(2, 0, device="cuda"),
                                          torch.full((5, 2), 10.0, device="cuda"), 100.0, return_error=True)
    assert float(err.abs().max()) == 0.0 and idx.cpu().tolist() == [0, 0] and best.cpu().tolist() == [10.0, 10.0]
    # identical images: zero photometric error at every hypothesis for a far object (disparity below the sampling resolution)
    uvz, w = da.sample(calib, 2, L.H, L.W, dev(box[:1] * 2), dev(poses[:1]), dev(borders[:1] * 2))
    de = torch.full((3, 1), 1e6, device="cuda")
    _, err, _ = da.enumeration_depth(L, L, uvz, w, de, float(calib.p2[0, 0] * 2 * 0.54), return_error=True)
    assert float(err.max()) < 1e-3 * float(w.sum())
