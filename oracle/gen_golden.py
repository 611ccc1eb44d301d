"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF.

Run in the build container (``/root/reference`` mounted):  ``python -m oracle.gen_golden``
The reference Python (stereo_network_old.py, decode.py, DCNv2/dcn_v2.py, DCNv2/test.py's KAT logic) is
imported unmodified through ``oracle/ref_loader.py`` (shims listed there); inputs are seeded; outputs are
stored as small fixtures.  Large bit-exact outputs are stored as sha256 + a strided sample.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote %-28s %8.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def gen_dcn(R):
    dcn = R.dcn
    # 1. the reference's own known-answer test: DCNv2/test.py:32-67 (check_zero_offset), N,C,H,W = 2,2,4,4
    torch.manual_seed(0)
    N, inC, inH, inW, outC, kH, kW = 2, 2, 4, 4, 2, 3, 3
    m = dcn.DCNv2(inC, outC, (kH, kW), stride=1, padding=1, dilation=1, deformable_groups=1)
    m.weight.data.zero_()
    m.bias.data.zero_()
    for p in range(inC):
        m.weight.data[p, p, kH // 2, kW // 2] = 1.0          # conv_identify, test.py:20-29
    x = torch.randn(N, inC, inH, inW)
    offset = torch.zeros(N, 2 * kH * kW, inH, inW)
    mask = torch.sigmoid(torch.zeros(N, kH * kW, inH, inW))
    out = m(x, offset, mask)
    assert (x - 2 * out).abs().max() < 1e-10                 # the reference's pass criterion
    save("dcn_kat_zero_offset", x=npy(x), offset=npy(offset), mask=npy(mask), weight=npy(m.weight), bias=npy(m.bias),
         y=npy(out))

    # 2. dcn_v2_conv forward + backward through the reference autograd Function (dcn_v2.py:16-51)
    cases = {"a": (2, 4, 6, 7, 3, 1, 1, 1, 1), "b": (1, 8, 9, 11, 5, 2, 1, 1, 1), "c": (2, 8, 7, 9, 4, 1, 2, 2, 2),
             "d": (1, 32, 12, 20, 16, 1, 1, 1, 1), "e": (1, 32, 8, 10, 8, 1, 1, 1, 2)}
    for tag, (B, Cin, H, W, Cout, stride, pad, dil, dg) in cases.items():
        torch.manual_seed(ord(tag))
        Ho = (H + 2 * pad - (dil * 2 + 1)) // stride + 1
        Wo = (W + 2 * pad - (dil * 2 + 1)) // stride + 1
        x = torch.randn(B, Cin, H, W, requires_grad=True)
        offset = (torch.randn(B, dg * 18, Ho, Wo) * 2).requires_grad_(True)      # ~N(0, 2^2) as test.py:74
        mask = torch.sigmoid(torch.randn(B, dg * 9, Ho, Wo)).requires_grad_(True)
        w = (torch.randn(Cout, Cin, 3, 3) * 0.2).requires_grad_(True)
        b = torch.rand(Cout, requires_grad=True)
        y = dcn.dcn_v2_conv(x, offset, mask, w, b, stride, pad, dil, dg)
        gy = torch.randn_like(y)
        y.backward(gy)
        save("dcn_conv_" + tag, x=npy(x), offset=npy(offset), mask=npy(mask), weight=npy(w), bias=npy(b), y=npy(y),
             gy=npy(gy), gx=npy(x.grad), goffset=npy(offset.grad), gmask=npy(mask.grad), gweight=npy(w.grad),
             gbias=npy(b.grad), cfg=np.array([stride, pad, dil, dg]))

    # 3. the DCN module (dcn_v2.py:97-128) with a non-zero conv_offset_mask
    torch.manual_seed(7)
    mod = dcn.DCN(16, 8, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1)
    mod.conv_offset_mask.weight.data.normal_(0, 0.08)
    mod.conv_offset_mask.bias.data.normal_(0, 0.2)
    mod.bias.data.uniform_(0, 1)
    x = torch.randn(2, 16, 10, 14)
    y = mod(x)
    save("dcn_module", x=npy(x), y=npy(y), **{"p_" + k: npy(v) for k, v in mod.state_dict().items()})


def gen_proposals_and_volume(R):
    net = R.net
    from side_b200.utils.synthetic import make_boxes
    torch.manual_seed(11)
    B, C, H, W = 2, 32, 24, 320          # full feature width so the 319 clamp is exercised
    # values rounded to fp16 so the fixture can store them exactly at half the size
    featL, featR = torch.randn(B, C, H, W).half().float(), torch.randn(B, C, H, W).half().float()
    left, right, _ = make_boxes(B, 3, seed=4, W4=320, H4=24)
    left[0, 1:] = torch.tensor([300.0, 2.0, 318.5, 9.0])     # hits the x clamp (:70-71)
    right[0, 1:] = torch.tensor([294.0, 2.5, 312.5, 9.5])
    left[1, 1:] = torch.tensor([1.0, 0.0, 9.0, 23.9])        # right box clamps at 0 (:75-76), touches the bottom
    right[1, 1:] = torch.tensor([-3.0, 0.0, 5.0, 23.9])
    left[2, 1:] = torch.tensor([100.0, 5.0, 100.0, 8.0])     # zero width: division by zero -> clamp to 87 (A.3)
    right[2, 1:] = torch.tensor([100.0, 5.0, 100.0, 8.0])
    fb = torch.tensor([384.38, 420.0])
    for D in (16, 48):
        pl, pr, db = net.get_proposal_shift(left, right, D, fb, None)
        save("proposal_shift_D%d" % D, left=npy(left), right=npy(right), fb=npy(fb), pro_left=npy(pl), pro_right=npy(pr),
             depth_bin=npy(db))
    # the volume build loop, stereo_network_old.py:366-376, with the reference's RoIAlign configuration (:271)
    from torchvision.ops import RoIAlign
    P = D = 16
    RoI = RoIAlign((P, P), spatial_scale=1, sampling_ratio=2)
    pl, pr, db = net.get_proposal_shift(left, right, D, fb, None)
    cost = torch.zeros(db.shape[0], C * 3, D, P, P)
    for ind in range(D):
        rl = RoI(featL, pl[ind])
        rr = RoI(featR, pr[ind])
        cost[:, :C, ind] = rl
        cost[:, C:2 * C, ind] = rr
        cost[:, 2 * C:, ind] = rl - rr
    # gate + aggregation + soft-argmin: the reference's cost_volume module itself (:135-244), seeded weights
    torch.manual_seed(5)
    est = net.cost_volume(64).eval()
    grabbed = {}
    est.dres0.register_forward_pre_hook(lambda m, inp: grabbed.__setitem__("gated", inp[0].detach().clone()))
    est.classify.register_forward_hook(lambda m, inp, out: grabbed.__setitem__("logits", out.detach().clone()))
    with torch.no_grad():
        disp = est(cost, D, db)
    gated = grabbed["gated"]
    num = (gated * cost).sum((1, 3, 4))
    den = (cost * cost).sum((1, 3, 4))
    xc = num / den                                            # per-(n,d) scalar the reference multiplied by
    cn = npy(cost)
    save("inst_costvol", featL=npy(featL.half()), featR=npy(featR.half()), left=npy(left), right=npy(right), fb=npy(fb),
         cost_sha256=np.array(sha(cn)), cost_sample=cn.reshape(-1)[::97].copy(), cost_shape=np.array(cn.shape),
         xcross=npy(xc), gated_sample=npy(gated).reshape(-1)[::97].copy(),
         logits=npy(grabbed["logits"]), depth_bin=npy(db), disp=npy(disp),
         est_state_sha256=np.array(sha(np.concatenate([npy(v).reshape(-1).astype(np.float64) for v in est.state_dict().values()]))))


def gen_decode(R):
    dec = R.decode
    torch.manual_seed(21)
    B, cat, H, W, grid, K = 2, 3, 24, 40, 4, 12
    hm = torch.randn(B, cat, H, W) * 1.5 - 2.19
    wh = torch.rand(B, 3, H, W) * 30 + 2
    reg = torch.rand(B, 3, H, W)
    reg[:, 1] -= 5.0
    # a row whose box coordinate sum 2*(cx+cy) is <= 0 gets dropped (decode.py:122-124): peak at (y=0, x=1)
    hm[1, 0, 0, 1] = 4.0
    reg[1, 0, 0, 1], reg[1, 2, 0, 1] = -3.0, -1.0
    kept = torch.randn(B, 6 * grid, H, W)
    dim = torch.randn(B, 3, H, W)
    orien = torch.randn(B, 2, H, W)
    bk, brk, shape = dec.bbox_decode(hm, wh, reg, K=K)
    heat = torch.sigmoid(hm)
    det, detr, info = dec.ddd_decode(heat.clone(), kept, dim, orien, wh, reg, grid, K=K)
    save("decode", hm=npy(hm), wh=npy(wh), reg=npy(reg), kept=npy(kept), dim=npy(dim), orien=npy(orien),
         bbox_keep=npy(bk), bbox_right_keep=npy(brk), bbox_shape=np.array(list(shape)), det=npy(det), det_right=npy(detr),
         info=npy(info), cfg=np.array([grid, K]))


def gen_e2e(R):
    from side_b200.networks import get_pose_net
    from side_b200.utils.synthetic import HEADS, make_batch, make_boxes, realistic_init
    torch.manual_seed(0)
    mine = realistic_init(get_pose_net(34, HEADS, 256), seed=1).eval()
    ref = R.net.get_pose_net(34, HEADS, 256).eval()
    ref.load_state_dict(mine.state_dict())
    state_sha = sha(np.concatenate([npy(v).reshape(-1).astype(np.float64) for v in mine.state_dict().values()]))
    batch = make_batch(1, 64, 1280, seed=3)
    left, right, shape = make_boxes(1, 6, seed=9, W4=320, H4=16)
    with torch.no_grad():
        z = ref(batch, True, None, 1.0)[0]
        zt = ref(batch, True, (left, right, shape), 1.0)[0]
        hm = z['hm'].clone().sigmoid_()
        det, detr, info = R.decode.ddd_decode(hm, z['kept_type'], z['dim'], z['orien'], wh=z['wh'], reg=z['reg'],
                                              grid_size=28, K=100)
    g = torch.Generator().manual_seed(2)
    pos = torch.randint(0, 16 * 320, (512,), generator=g)
    pick = lambda t: npy(t.reshape(t.shape[0], t.shape[1], -1)[:, :, pos])
    save("e2e_small", state_sha256=np.array(state_sha), input_sha256=np.array(sha(npy(batch['input']))),
         hm=npy(z['hm']), wh=npy(z['wh']), reg=npy(z['reg']), depth=npy(z['depth']), depth_target=npy(zt['depth']),
         pos=npy(pos), kept_type_s=pick(z['kept_type']), dim_s=pick(z['dim']), orien_s=pick(z['orien']),
         det=npy(det), det_right=npy(detr), info=npy(info))


def dense_align_case():
    """Seeded inputs of the dense-alignment fixture (shared with the tests): a 96 x 320 textured stereo pair whose right view
    is the left one shifted by the disparity of ~20 m, KITTI-like calibration at 1/4 scale, 4 RoIs (one of them with no
    valid pixel: its box lies beside the projected cuboid)."""
    import types
    rng = np.random.RandomState(5)
    H, W = 96, 320
    base = rng.rand(H // 4 + 2, W // 4 + 2, 3).astype(np.float32)
    big = np.kron(base, np.ones((4, 4, 1), np.float32))[:H + 8, :W + 8]
    k = np.ones(5, np.float32) / 5                                    # box blur -> smooth texture with gradients
    for ax in (0, 1):
        big = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), ax, big)
    img_l = (np.clip(big[4:4 + H, 4:4 + W], 0, 1) * 255).astype(np.uint8)
    img_r = np.ascontiguousarray(np.roll(img_l, -4, axis=1))
    p2 = np.array([[721.54, 0, 609.56, 44.86], [0, 721.54, 172.85, 0.216], [0, 0, 1, 0.00275]], np.float32)
    p2[:2] /= 4
    p3 = p2.copy()
    p3[0, 3] = -339.52 / 4
    calib = types.SimpleNamespace(p2=p2, p3=p3)
    opt = types.SimpleNamespace(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    box = np.array([[100., 30., 160., 70.], [200., 40., 230., 60.], [120., 20., 200., 80.], [20., 35., 70., 66.]], np.float32)
    borders = np.array([[105., 150.], [202., 228.], [125., 190.], [24., 66.]], np.float32)
    poses = np.array([[-1.5, 1.6, 20., 1.6, 1.5, 3.9, 0.3], [2.0, 1.7, 30., 1.6, 1.5, 3.9, -1.2],
                      [0.2, 1.5, 8., 1.8, 1.6, 4.2, 1.3], [-9.5, 1.4, 18., 1.7, 1.5, 4.0, -0.4]], np.float32)
    return img_l, img_r, calib, opt, box, borders, poses


def gen_dense_align(R):
    """F2: the reference's sample / enumeration_depth / align_parallel (dense_align/dense_align.py) executed unmodified."""
    import importlib
    import warnings
    import torch.nn.functional as F
    da = importlib.import_module("dense_align.dense_align")
    img_l, img_r, calib, opt, box, borders, poses = dense_align_case()
    box_t, borders_t, poses_t = torch.from_numpy(box), torch.from_numpy(borders), torch.from_numpy(poses)
    H, W = img_l.shape[:2]
    scale = 2
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        uvz, wgt = da.sample(calib, scale, 2 * H, 2 * W, box_t * scale, poses_t, borders_t * scale)
        status, best_dis = da.align_parallel(calib, opt, img_l, img_r, box_t, borders_t, poses_t)
        # the two images exactly as align_parallel prepares them (:251-266), and one enumeration with its error matrix
        mean = np.array(opt.mean, dtype=np.float32).reshape(1, 1, 3)
        std = np.array(opt.std, dtype=np.float32).reshape(1, 1, 3)

        def prep(im):
            im = (im.astype(np.float32) / 255.)
            im = ((im - mean) / std).transpose(2, 0, 1)[np.newaxis, ...]
            return F.interpolate(torch.from_numpy(im), scale_factor=2, mode='bilinear', align_corners=False)
        im_l, im_r = prep(img_l), prep(img_r)
        f = calib.p2[0, 0] * scale
        bl = (calib.p2[0, 3] - calib.p3[0, 3]) * scale / f
        dis_init = f * bl / poses_t[:, 2]
        depth_enum = torch.zeros(50, box_t.size(0))
        for i in range(50):
            depth_enum[i] = dis_init.reciprocal() * f * bl - 50 * 0.5 / 2 + 0.5 * i
        depth_enum[depth_enum < 1.5] = 1.5
        seen = {}
        tmin = torch.min

        def spy(*a, **k):                       # enumeration_depth only returns the winner: keep the error matrix it minimises
            seen["err"] = a[0].clone()
            return tmin(*a, **k)
        torch.min = spy
        try:
            best = da.enumeration_depth(im_l, im_r, uvz, wgt, depth_enum, f * bl)
        finally:
            torch.min = tmin
    pos = np.random.RandomState(1).randint(0, 4 * H * W, 4096)
    save("dense_align", img_l=img_l, img_r=img_r, uvz=npy(uvz), weight=npy(wgt), status=npy(status), best_dis=npy(best_dis),
         depth_enum=npy(depth_enum), err_sum=npy(seen["err"]), best_depth=npy(best), fb=np.float32(f * bl),
         im_l_sha256=np.array(sha(npy(im_l))), im_pos=pos, im_l_s=npy(im_l).reshape(3, -1)[:, pos],
         im_r_s=npy(im_r).reshape(3, -1)[:, pos])


def voxel_new_case():
    """Seeded inputs of the stereo_network_new fixture (numpy only; shared with the tests).  A quarter-size configuration
    (network input 96 x 320, features 24 x 80) keeps the fixture small: the reference module's size constants
    ``input_h, input_w`` (stereo_network_new.py:20) are set to 96, 320 for the run -- a configuration value, no code changes."""
    from side_b200.preprocess import get_affine_transform
    H_in, W_in = 96, 320
    p2 = np.array([[721.54, 0, 609.56, 44.86], [0, 721.54, 172.85, 0.216], [0, 0, 1, 0.00275]], np.float32)
    p3 = p2.copy()
    p3[0, 3] = -339.52
    c, s = np.array([621., 187.5], np.float32), np.array([1242, 375], np.int32)
    trans = get_affine_transform(c, s, 0, [W_in // 4, H_in // 4]).astype(np.float32)
    trans_inv = get_affine_transform(c, s, 0, [W_in // 4, H_in // 4], inv=1).astype(np.float32)
    B = 2
    fb = np.full((B,), 721.54 * 0.5327, np.float32)
    # boxes in feature pixels (80 x 24): (image, x1, y1, x2, y2); right box = left shifted by the disparity of 12 .. 40 m
    left = np.array([[0, 30.0, 10.0, 36.0, 14.0], [0, 52.5, 11.0, 56.0, 13.5], [0, 2.0, 9.0, 9.5, 15.0],
                     [1, 41.0, 10.5, 45.0, 13.0], [1, 70.0, 9.5, 78.5, 15.5]], np.float32)
    disp = np.array([1.9, 0.9, 2.0, 1.2, 0.62], np.float32)
    right = left.copy()
    right[:, 1] -= disp
    right[:, 3] -= disp
    stack = lambda a: np.ascontiguousarray(np.broadcast_to(a, (B,) + a.shape)).astype(np.float32)
    return dict(H_in=H_in, W_in=W_in, p2=stack(p2), p3=stack(p3), trans=stack(trans), trans_inv=stack(trans_inv), fb=fb,
                left=left, right=right)


def gen_voxel_new(R):
    """F3: the reference's get_proposal_shift / get_voxel executed directly, and stereo_network.forward of the voxel variant
    executed with ground-truth RoIs; the voxel tensor is captured at the input of self.pointNet, the reduced features at
    the output of self.feaRuduce (forward hooks: no reference code is changed)."""
    import importlib
    import warnings
    rn = importlib.import_module("models.networks.stereo_network_new")
    c = voxel_new_case()
    rn.input_h, rn.input_w = float(c["H_in"]), float(c["W_in"])
    t = lambda k: torch.from_numpy(c[k])
    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore")
        pro_l, pro_r, depth_bin = rn.get_proposal_shift(t("left"), t("right"), 20, t("fb"), t("trans_inv"))
        vox = rn.get_voxel(t("left"), t("right"), t("p2"), t("p3"), t("fb"), depth_bin, t("trans"), t("trans_inv"))
        torch.manual_seed(7)
        heads = {'hm': 3, 'wh': 3, 'reg': 3, 'dim': 3, 'orien': 2, 'kept_type': 168}
        net = rn.get_pose_net(34, heads, 256).eval()
        seen = {"fea": []}
        net.feaRuduce.register_forward_hook(lambda m, i, o: seen["fea"].append(o.detach().clone()))
        net.pointNet.register_forward_hook(lambda m, i, o: seen.update(voxel=i[0].detach().clone(), disp=o.detach().clone()))
        g = torch.Generator().manual_seed(11)
        batch = {'input': torch.randn(2, 3, c["H_in"], c["W_in"], generator=g), 'input_right': torch.randn(2, 3, c["H_in"], c["W_in"], generator=g),
                 'fb': t("fb"), 'p2': t("p2"), 'p3': t("p3"), 'trans': t("trans"), 'trans_inv': t("trans_inv")}
        z = net(batch, True, (t("left"), t("right"), torch.Size([2, 50, 1])))[0]
    voxel = npy(seen["voxel"])                                       # [5, 192, 1000]
    pos = np.random.RandomState(2).randint(0, voxel.size, 20000)
    save("voxel_new", pro_left=npy(pro_l), pro_right=npy(pro_r), depth_bin=npy(depth_bin),
         norm3=npy(vox[0]), valid3=npy(vox[1]), normL=npy(vox[2]), validL=npy(vox[3]), normR=npy(vox[4]), validR=npy(vox[5]),
         depth_ori=npy(vox[6]), feaL=npy(seen["fea"][0]), feaR=npy(seen["fea"][1]), voxel_sha256=np.array(sha(voxel)),
         voxel_pos=pos, voxel_s=voxel.reshape(-1)[pos], voxel_absmax=np.float32(np.abs(voxel).max()),
         voxel_nonzero=np.int64((voxel != 0).sum()), disp=npy(seen["disp"]), depth=npy(z["depth"]))


def main():
    os.makedirs(OUT, exist_ok=True)
    R = ref_loader.load()
    gen_dcn(R)
    gen_proposals_and_volume(R)
    gen_decode(R)
    gen_e2e(R)
    gen_dense_align(R)
    gen_voxel_new(R)


if __name__ == "__main__":
    main()
