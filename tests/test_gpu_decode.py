"""GPU parity: single-launch NMS + top-K + gather decode vs the oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

from oracle import c_oracle as co  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_decode_golden(lib):
    from side_b200.decode import bbox_decode, ddd_decode
    g = golden("decode")
    grid, K = [int(v) for v in g["cfg"]]
    bk, brk, shape = bbox_decode(dev(g["hm"]), dev(g["wh"]), dev(g["reg"]), K=K)
    assert list(shape) == list(g["bbox_shape"])
    assert bk.shape == g["bbox_keep"].shape, "row dropping (decode.py:122-124) differs"
    assert np.abs(bk.cpu().numpy() - g["bbox_keep"]).max() < 1e-4
    assert np.abs(brk.cpu().numpy() - g["bbox_right_keep"]).max() < 1e-4
    heat = torch.sigmoid(dev(g["hm"]))
    det, detr, info = ddd_decode(heat, dev(g["kept"]), dev(g["dim"]), dev(g["orien"]), dev(g["wh"]), dev(g["reg"]), grid, K=K)
    ref_info = g["info"].copy()
    ref_info[..., 8] = np.floor(ref_info[..., 8])
    assert np.abs(det.cpu().numpy() - g["det"]).max() < 1e-5
    assert np.abs(detr.cpu().numpy() - g["det_right"]).max() < 1e-5
    assert np.abs(info.cpu().numpy() - ref_info).max() < 1e-6
    assert np.array_equal(det.cpu().numpy()[..., 5], g["det"][..., 5])


@pytest.mark.parametrize("cfg", [(1, 3, 96, 320, 100), (4, 3, 96, 320, 100), (2, 1, 13, 17, 5), (3, 5, 8, 8, 64), (1, 2, 4, 4, 16), (2, 3, 24, 40, 1)])
def test_topk_indices_bit_exact_vs_oracle(lib, cfg):
    from side_b200 import ops
    B, Cat, H, W, K = cfg
    rng = np.random.default_rng(B + Cat + H)
    hm = (rng.standard_normal((B, Cat, H, W)) * 1.5 - 2.19).astype(np.float32)
    wh = (rng.random((B, 3, H, W)) * 30 + 2).astype(np.float32)
    reg = rng.random((B, 3, H, W)).astype(np.float32)
    o = ops.bbox_decode_raw(dev(hm), dev(wh), dev(reg), K=K, wh_scale=1.3)
    heat = torch.sigmoid(dev(hm)).cpu().numpy()              # same sigmoid values for both sides of the index check
    score, ind, cls, ys, xs = co.nms_topk(heat, K, heat_is_logit=False)
    o2 = ops.bbox_decode_raw(dev(heat), dev(wh), dev(reg), K=K, wh_scale=1.3, heat_is_logit=False)
    assert np.array_equal(o2["ind"].cpu().numpy(), ind) and np.array_equal(o2["cls"].cpu().numpy(), cls)
    assert np.array_equal(o2["score"].cpu().numpy(), score)
    bbox, bbr, keep = co.bbox_decode(hm, wh, reg, K, wh_scale=1.3)
    # fused-sigmoid flavour: indices agree wherever the oracle's own expf-based sigmoid has no near-ties
    assert np.mean(o["ind"].cpu().numpy() == ind) > 0.99
    same = (o["ind"].cpu().numpy() == ind)
    assert np.abs(o["bbox"].cpu().numpy() - bbox)[same].max() < 1e-4
    k = o["keep"].cpu().numpy().reshape(B, K).astype(bool)
    slot = o["slot"].cpu().numpy()
    assert np.array_equal(o["count"].cpu().numpy(), k.sum(1))
    assert np.array_equal(slot[k], (np.cumsum(k, 1) - 1)[k])


def test_ties_resolve_to_lowest_index(lib):
    """All-equal and all-zero maps: documented tie rule."""
    from side_b200 import ops
    heat = torch.zeros(1, 2, 6, 7, device="cuda")
    z = torch.zeros(1, 3, 6, 7, device="cuda")
    o = ops.bbox_decode_raw(heat, z, z, K=5, heat_is_logit=False)
    assert o["ind"].cpu().tolist() == [[0, 1, 2, 3, 4]] and o["cls"].cpu().tolist() == [[0, 0, 0, 0, 0]]
    heat[0, 1, 3, 3] = 0.7
    heat[0, 0, 5, 6] = 0.7
    o = ops.bbox_decode_raw(heat, z, z, K=3, heat_is_logit=False)
    assert o["ind"].cpu().tolist() == [[41, 24, 0]] and o["cls"].cpu().tolist() == [[0, 1, 0]]
    sc, ind, cls, _, _ = co.nms_topk(heat.cpu().numpy(), 3)
    assert ind.tolist() == [[41, 24, 0]] and cls.tolist() == [[0, 1, 0]]


def test_plateau_nms_keeps_all_equal_neighbours(lib):
    """keep = (maxpool == heat) keeps every pixel of a flat maximum (decode.py:14)."""
    from side_b200 import ops
    heat = torch.full((1, 1, 5, 5), 0.1, device="cuda")
    heat[0, 0, 1:3, 1:3] = 0.9
    z = torch.zeros(1, 3, 5, 5, device="cuda")
    o = ops.bbox_decode_raw(heat, z, z, K=6, heat_is_logit=False)
    assert o["ind"].cpu().tolist()[0][:4] == [6, 7, 11, 12]
    assert o["score"].cpu().tolist()[0][4] == 0.0 or abs(o["score"].cpu().tolist()[0][4] - 0.1) < 1e-7


def test_ddd_decode_full_size_vs_oracle(lib):
    from side_b200 import ops
    rng = np.random.default_rng(3)
    B, H, W, grid, K = 2, 96, 320, 28, 100
    heat = (1 / (1 + np.exp(-(rng.standard_normal((B, 3, H, W)) * 1.5 - 2.19)))).astype(np.float32)
    kept = rng.standard_normal((B, 6 * grid, H, W)).astype(np.float32)
    dim = rng.standard_normal((B, 3, H, W)).astype(np.float32)
    orien = rng.standard_normal((B, 2, H, W)).astype(np.float32)
    wh = rng.random((B, 3, H, W)).astype(np.float32) * 30
    reg = rng.random((B, 3, H, W)).astype(np.float32)
    det, detr, info = ops.ddd_decode_raw(dev(heat), dev(kept), dev(dim), dev(orien), dev(wh), dev(reg), grid, K=K)
    rd, rdr, ri = co.ddd_decode(heat, kept, dim, orien, wh, reg, grid, K)
    assert np.array_equal(det.cpu().numpy(), rd) and np.array_equal(detr.cpu().numpy(), rdr)
    assert np.array_equal(info.cpu().numpy(), ri)


def test_small_map_in_a_fresh_process(lib):
    """Regression: with a small H*W the kernel's static (33 KB) + dynamic shared memory exceeds the 48 KB default although the
    dynamic part alone does not; in a fresh process (no earlier large launch to raise the attribute) the launch used to fail."""
    import subprocess, sys, os
    code = ("import sys; sys.path.insert(0, %r); import torch; from side_b200 import ops; "
            "h = torch.randn(1, 3, 16, 320, device='cuda'); o = ops.bbox_decode_raw(h, torch.rand_like(h), torch.rand_like(h), K=100); "
            "torch.cuda.synchronize(); print(int(o['count'].sum()))") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_heat_map_too_large_for_shared_memory_is_refused(lib):
    """static (~33 KB) + dynamic shared memory are checked together -- a 200x260 map (208 KB of keys) must come back
    as SIDE_ERR_UNSUPPORTED with a message, not as an opaque launch failure."""
    from side_b200 import ops
    hm = torch.randn(1, 1, 200, 260, device="cuda")
    wh = torch.rand(1, 3, 200, 260, device="cuda")
    with pytest.raises(RuntimeError, match="shared memory"):
        ops.bbox_decode_raw(hm, wh, wh, K=10)
    torch.cuda.synchronize()                                  # the context is still healthy
    o = ops.bbox_decode_raw(hm[:, :, :96, :].contiguous(), wh[:, :, :96, :].contiguous(), wh[:, :, :96, :].contiguous(), K=10)
    assert o["score"].shape == (1, 10)
