"""CPU: the oracle (oracle/side_oracle.c + oracle/torch_port.py) against the golden vectors that were produced by
executing the reference itself (oracle/gen_golden.py).  This is what pins the oracle."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden, rel_err
from oracle import c_oracle as co
from oracle import torch_port as tp


def test_dcn_zero_offset_kat():
    """The reference's own KAT, DCNv2/test.py:32-67: identity-centre weight, zero offsets, mask 0.5 => 2*y == x."""
    g = golden("dcn_kat_zero_offset")
    y = co.dcn_forward(g["x"], g["offset"], g["mask"], g["weight"], g["bias"])
    assert np.abs(g["x"] - 2 * y).max() < 1e-10
    assert np.array_equal(y, g["y"])


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e"])
def test_dcn_forward_backward_vs_reference(tag):
    g = golden("dcn_conv_" + tag)
    stride, pad, dil, dg = [int(v) for v in g["cfg"]]
    y = co.dcn_forward(g["x"], g["offset"], g["mask"], g["weight"], g["bias"], stride, pad, dil, dg)
    assert rel_err(y, g["y"]) < 1e-5
    gx, go, gm, gw, gb = co.dcn_backward(g["x"], g["offset"], g["mask"], g["weight"], g["gy"], stride, pad, dil, dg)
    for mine, ref in ((gx, "gx"), (go, "goffset"), (gm, "gmask"), (gw, "gweight"), (gb, "gbias")):
        assert rel_err(mine, g[ref]) < 1e-5, ref


def test_dcn_module_port():
    g = golden("dcn_module")
    om = torch.nn.functional.conv2d(torch.from_numpy(g["x"]), torch.from_numpy(g["p_conv_offset_mask.weight"]),
                                    torch.from_numpy(g["p_conv_offset_mask.bias"]), padding=1)
    y = tp.dcn_module_forward(torch.from_numpy(g["x"]), om, torch.from_numpy(g["p_weight"]), torch.from_numpy(g["p_bias"]),
                              1, 1, 1)
    assert rel_err(y.numpy(), g["y"]) < 1e-6
    # C oracle on the same offsets / sigmoid(mask)
    o = om.numpy()
    yc = co.dcn_forward(g["x"], o[:, :18], 1.0 / (1.0 + np.exp(-o[:, 18:].astype(np.float64))).astype(np.float32),
                        g["p_weight"], g["p_bias"])
    assert rel_err(yc, g["y"]) < 1e-5


@pytest.mark.parametrize("D", [16, 48])
def test_proposal_shift_bit_exact(D):
    g = golden("proposal_shift_D%d" % D)
    pl, pr, db = co.proposal_shift(g["left"], g["right"], g["fb"], D)
    assert np.array_equal(db.view(np.uint32), g["depth_bin"].view(np.uint32))
    assert np.array_equal(pl.view(np.uint32), g["pro_left"].view(np.uint32))
    assert np.array_equal(pr.view(np.uint32), g["pro_right"].view(np.uint32))
    # edge cases present in the fixture: x clamp at 319, right clamp at 0, zero-width box -> depth_min 87
    assert g["pro_left"][:, 0, 3].max() == 319.0 and g["pro_right"][:, 1, 1].min() == 0.0
    assert np.all(g["depth_bin"][2] == 87.0)
    tl, tr, tdb = tp.proposal_shift(torch.from_numpy(g["left"]), torch.from_numpy(g["right"]), torch.from_numpy(g["fb"]), D, 319.)
    assert np.array_equal(tdb.numpy(), g["depth_bin"]) and np.array_equal(tl.numpy(), g["pro_left"])


def test_inst_costvol_bit_exact_and_gate_and_softargmin():
    g = golden("inst_costvol")
    fL, fR = g["featL"].astype(np.float32), g["featR"].astype(np.float32)
    D = P = 16
    pl, pr, db = co.proposal_shift(g["left"], g["right"], g["fb"], D)
    cost = co.inst_costvol(fL, fR, pl, pr, P)
    assert list(cost.shape) == list(g["cost_shape"])
    assert hashlib.sha256(cost.tobytes()).hexdigest() == str(g["cost_sha256"])      # bit-exact vs the reference loop
    assert np.array_equal(cost.reshape(-1)[::97], g["cost_sample"])
    gated, xc = co.xcross_gate(cost, 32)
    assert rel_err(xc, g["xcross"]) < 1e-4
    assert rel_err(gated.reshape(-1)[::97], g["gated_sample"]) < 1e-5
    depth, _ = co.softargmin(g["logits"][:, 0], g["depth_bin"])
    assert rel_err(depth, g["disp"]) < 1e-5
    # torch port of the same sequence
    c2, db2 = tp.inst_costvol(torch.from_numpy(fL), torch.from_numpy(fR), torch.from_numpy(g["left"]),
                              torch.from_numpy(g["right"]), torch.from_numpy(g["fb"]), D, P, 319.)
    assert np.array_equal(c2.numpy(), cost) and np.array_equal(db2.numpy(), db)
    d2 = tp.softargmin(torch.from_numpy(g["logits"][:, 0]), torch.from_numpy(g["depth_bin"]))
    assert rel_err(d2.numpy(), g["disp"]) < 1e-6


def test_decode_vs_reference():
    g = golden("decode")
    grid, K = [int(v) for v in g["cfg"]]
    bbox, bbr, keep = co.bbox_decode(g["hm"], g["wh"], g["reg"], K)
    assert not keep.all()                                   # the fixture drops rows (decode.py:122-124)
    bk, brk = bbox.reshape(-1, 5)[keep], bbr.reshape(-1, 5)[keep]
    assert bk.shape == g["bbox_keep"].shape
    assert np.abs(bk - g["bbox_keep"]).max() < 1e-4 and np.abs(brk - g["bbox_right_keep"]).max() < 1e-4
    assert np.array_equal(bk[:, 0], g["bbox_keep"][:, 0])
    heat = (1.0 / (1.0 + np.exp(-g["hm"].astype(np.float64)))).astype(np.float32)
    det, detr, info = co.ddd_decode(heat, g["kept"], g["dim"], g["orien"], g["wh"], g["reg"], grid, K)
    assert np.abs(det - g["det"]).max() < 1e-5 and np.abs(detr - g["det_right"]).max() < 1e-5
    ref_info = g["info"].copy()
    ref_info[..., 8] = np.floor(ref_info[..., 8])           # torch>=1.5 true division in the reference
    assert np.abs(info - ref_info).max() < 1e-6
    # class ids and integer pixel centres are exact
    assert np.array_equal(det[..., 5], g["det"][..., 5])


def test_volume_builders_against_torch_restatement():
    """PARITY UNPINNED by the reference (no call site): C oracle vs the torch restatement only."""
    rng = np.random.default_rng(0)
    L = rng.standard_normal((2, 8, 5, 24)).astype(np.float32)
    R = rng.standard_normal((2, 8, 5, 24)).astype(np.float32)
    v = co.concat_volume(L, R, 6)
    assert np.array_equal(v, tp.concat_volume(torch.from_numpy(L), torch.from_numpy(R), 6).numpy())
    assert np.array_equal(v[:, :8, 0], L) and np.array_equal(v[:, 8:, 0], R)     # D=1 slice == torch.cat((L,R),1) (:348)
    w = co.gwc_volume(L, R, 6, 4)
    assert rel_err(w, tp.gwc_volume(torch.from_numpy(L), torch.from_numpy(R), 6, 4).numpy()) < 1e-6


def test_conv3d_layer_restatement_vs_torch():
    """The aggregation-network layer of the oracle (conv3d 3x3x3 pad 1 + folded BatchNorm3d + ReLU + residual, MaxPool3d(1,2,2))
    against the ops the reference module calls (nn.Conv3d / BatchNorm3d / MaxPool3d, stereo_network_old.py:139-171)."""
    import torch
    import torch.nn.functional as F
    from oracle import c_oracle as co
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 6, 4, 5, 7, generator=g)
    w = torch.randn(8, 6, 3, 3, 3, generator=g) * 0.2
    bn = torch.nn.BatchNorm3d(8).eval()
    bn.running_mean.normal_(0, 0.2, generator=g); bn.running_var.uniform_(0.5, 1.5, generator=g)
    bn.weight.data.uniform_(0.8, 1.2, generator=g); bn.bias.data.normal_(0, 0.1, generator=g)
    res = torch.randn(2, 8, 4, 5, 7, generator=g)
    with torch.no_grad():
        ref = F.relu(bn(F.conv3d(x, w, padding=1))) + res
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach()
    shift = (bn.bias - bn.running_mean * scale).detach()
    out = co.conv3d_bn_relu(x.numpy(), w.numpy(), scale.numpy(), shift.numpy(), relu=True, residual=res.numpy())
    assert rel_err(out, ref.numpy()) < 1e-5
    y = torch.randn(1, 3, 2, 6, 8, generator=g)
    assert np.array_equal(co.maxpool_hw2(y.numpy()), F.max_pool3d(y, (1, 2, 2)).numpy())


# ---- F2: dense photometric alignment (dense_align/dense_align.py), golden = the reference executed on the seeded case ----
def _da_case():
    from oracle.gen_golden import dense_align_case
    return dense_align_case()


def test_dense_align_sample_bit_exact():
    g = golden("dense_align")
    img_l, _, calib, _, box, borders, poses = _da_case()
    H, W = img_l.shape[:2]
    f, cx, cy = calib.p2[0, 0] * 2, calib.p2[0, 2] * 2, calib.p2[1, 2] * 2
    uvz, wgt, cnt = co.da_sample(box * 2, borders * 2, poses, float(f), float(cx), float(cy), 2 * H, 2 * W)
    assert np.array_equal(wgt, g["weight"])                 # same pixels kept, same row-major order, same padding
    assert np.array_equal(uvz, g["uvz"])
    assert list(cnt) == [int(v) for v in g["weight"].sum(1)] and cnt[1] == 0


def test_dense_align_prep_and_enumeration():
    g = golden("dense_align")
    img_l, img_r, _, opt, _, _, _ = _da_case()
    assert np.array_equal(img_l, g["img_l"]) and np.array_equal(img_r, g["img_r"])
    L, R = co.da_prep_u8(img_l, opt.mean, opt.std), co.da_prep_u8(img_r, opt.mean, opt.std)
    assert np.abs(L.reshape(3, -1)[:, g["im_pos"]] - g["im_l_s"]).max() < 2e-6          # ATen's interpolate: <= 2 ulp
    assert np.abs(R.reshape(3, -1)[:, g["im_pos"]] - g["im_r_s"]).max() < 2e-6
    err, best, idx = co.da_enum(L, R, g["uvz"], g["weight"], g["depth_enum"], float(g["fb"]))
    assert rel_err(err, g["err_sum"]) < 1e-5
    assert np.array_equal(best, g["best_depth"])
    assert np.array_equal(idx, g["err_sum"].argmin(0))


# ---- F3: stereo_network_new voxel volume; golden = the reference's get_voxel / forward executed on the seeded case ----
def test_voxel_new_coords_and_volume_vs_reference():
    from oracle.gen_golden import voxel_new_case
    g, c = golden("voxel_new"), voxel_new_case()
    out = co.voxel_coords(c["left"], c["right"], c["p2"], c["p3"], c["fb"], c["trans"], c["trans_inv"], g["depth_bin"],
                          c["H_in"], c["W_in"])
    for mine, key in zip(out, ("norm3", "valid3", "normL", "validL", "normR", "validR", "depth_ori")):
        if key.startswith("valid"):
            assert np.array_equal(mine, g[key]), key
        else:
            assert np.abs(mine - g[key]).max() < 1e-5, key            # torch.mm's summation order: measured 4e-6
    voxel, dori = co.voxel_volume(g["feaL"], g["feaR"], c["left"], c["right"], c["p2"], c["p3"], c["fb"], c["trans"],
                                  c["trans_inv"], c["H_in"], c["W_in"])
    assert np.abs(dori - g["depth_ori"]).max() < 1e-5 * np.abs(g["depth_ori"]).max()
    d = np.abs(voxel.reshape(-1)[g["voxel_pos"]] - g["voxel_s"])
    assert d.max() < 1e-5 * float(g["voxel_absmax"])                # measured 1e-8 absolute
    assert abs(int((voxel != 0).sum()) - int(g["voxel_nonzero"])) < 1e-4 * voxel.size


def test_voxel_new_torch_port_and_proposals():
    from oracle.gen_golden import voxel_new_case
    from side_b200.networks import stereo_network_new as sn
    g, c = golden("voxel_new"), voxel_new_case()
    t = lambda k: torch.from_numpy(c[k])
    pl, pr, db = sn.get_proposal_shift(t("left"), t("right"), 20, t("fb"), t("trans_inv"))
    assert np.abs(db.numpy() - g["depth_bin"]).max() < 1e-5 * 90
    assert np.abs(pl.numpy() - g["pro_left"]).max() < 1e-5 * 80 and np.abs(pr.numpy() - g["pro_right"]).max() < 1e-5 * 80
    voxel, dori = tp.voxel_volume(torch.from_numpy(g["feaL"]), torch.from_numpy(g["feaR"]), t("left"), t("right"), t("p2"), t("p3"),
                                  t("fb"), t("trans"), t("trans_inv"), c["H_in"], c["W_in"])
    d = np.abs(voxel.numpy().reshape(-1)[g["voxel_pos"]] - g["voxel_s"])
    assert (d > 1e-4 * float(g["voxel_absmax"])).mean() < 2e-3
    assert np.abs(dori.numpy() - g["depth_ori"]).max() < 1e-5 * 40
