"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front-end of ``oracle/side_oracle.c``.

Every function takes / returns C-contiguous float32 numpy arrays and mirrors one reference
function (see the file:line citations in ``side_oracle.c``).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libside_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "side_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libside_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _chk(rc, name):
    if rc < 0:
        raise RuntimeError("%s failed with code %d" % (name, rc))


def dcn_forward(x, offset, mask, w, bias, stride=1, pad=1, dil=1, dg=1):
    x, px = _f(x); offset, po = _f(offset); mask, pm = _f(mask); w, pw = _f(w); bias, pb = _f(bias)
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    y = np.empty((B, Cout, Ho, Wo), np.float32)
    _chk(lib().orc_dcn_forward(px, po, pm, pw, pb, y.ctypes.data_as(C.POINTER(C.c_float)), B, Cin, H, W, Cout,
                               kh, kw, stride, stride, pad, pad, dil, dil, dg), "orc_dcn_forward")
    return y


def dcn_backward(x, offset, mask, w, gy, stride=1, pad=1, dil=1, dg=1):
    x, px = _f(x); offset, po = _f(offset); mask, pm = _f(mask); w, pw = _f(w); gy, pg = _f(gy)
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    gx = np.empty_like(x); go = np.empty_like(offset); gm = np.empty_like(mask); gw = np.empty_like(w)
    gb = np.empty((Cout,), np.float32)
    P = C.POINTER(C.c_float)
    _chk(lib().orc_dcn_backward(px, po, pm, pw, pg, gx.ctypes.data_as(P), go.ctypes.data_as(P),
                                gm.ctypes.data_as(P), gw.ctypes.data_as(P), gb.ctypes.data_as(P), B, Cin, H, W,
                                Cout, kh, kw, stride, stride, pad, pad, dil, dil, dg), "orc_dcn_backward")
    return gx, go, gm, gw, gb


def roi_align(feat, rois, P, spatial_scale=1.0, sampling_ratio=2):
    feat, pf = _f(feat); rois, pr = _f(rois)
    B, Cc, H, W = feat.shape
    N = rois.shape[0]
    out = np.empty((N, Cc, P, P), np.float32)
    _chk(lib().orc_roi_align(pf, pr, out.ctypes.data_as(C.POINTER(C.c_float)), N, Cc, H, W, P,
                             C.c_float(spatial_scale), sampling_ratio), "orc_roi_align")
    return out


def proposal_shift(left, right, fbs, D, x_clamp=319.0):
    """Boxes must be grouped by image (ascending b).  Returns pro_left[D,N,5], pro_right[D,N,5], depth_bin[N,D]."""
    left, pl = _f(left); right, pr = _f(right); fbs, pfb = _f(fbs)
    N = left.shape[0]
    a = np.empty((D, N, 5), np.float32); b = np.empty((D, N, 5), np.float32); db = np.empty((N, D), np.float32)
    P = C.POINTER(C.c_float)
    _chk(lib().orc_proposal_shift(pl, pr, pfb, N, D, C.c_float(x_clamp), a.ctypes.data_as(P), b.ctypes.data_as(P),
                                  db.ctypes.data_as(P)), "orc_proposal_shift")
    return a, b, db


def inst_costvol(featL, featR, pro_left, pro_right, P, sampling_ratio=2):
    featL, pL = _f(featL); featR, pR = _f(featR); pro_left, pl = _f(pro_left); pro_right, pr = _f(pro_right)
    B, Cc, H, W = featL.shape
    D, N, _ = pro_left.shape
    cost = np.empty((N, 3 * Cc, D, P, P), np.float32)
    _chk(lib().orc_inst_costvol(pL, pR, pl, pr, cost.ctypes.data_as(C.POINTER(C.c_float)), N, Cc, H, W, D, P,
                                sampling_ratio), "orc_inst_costvol")
    return cost


def xcross_gate(cost, Cc):
    cost, pc = _f(cost)
    N, C3, D, P, _ = cost.shape
    assert C3 == 3 * Cc
    out = np.empty_like(cost); xc = np.empty((N, D), np.float32)
    Pf = C.POINTER(C.c_float)
    _chk(lib().orc_xcross_gate(pc, out.ctypes.data_as(Pf), xc.ctypes.data_as(Pf), N, Cc, D, P), "orc_xcross_gate")
    return out, xc


def softargmin(logits, depth_bin):
    """logits [N,D,S,S] -> depth [N], prob [N,D]"""
    logits, pl = _f(logits); depth_bin, pd = _f(depth_bin)
    N, D, S, _ = logits.shape
    depth = np.empty((N,), np.float32); prob = np.empty((N, D), np.float32)
    Pf = C.POINTER(C.c_float)
    _chk(lib().orc_softargmin(pl, pd, depth.ctypes.data_as(Pf), prob.ctypes.data_as(Pf), N, D, S), "orc_softargmin")
    return depth, prob


def nms_topk(heat, K, heat_is_logit=False):
    heat, ph = _f(heat)
    B, Cat, H, W = heat.shape
    score = np.empty((B, K), np.float32); ys = np.empty((B, K), np.float32); xs = np.empty((B, K), np.float32)
    ind = np.empty((B, K), np.int32); cls = np.empty((B, K), np.int32)
    Pf, Pi = C.POINTER(C.c_float), C.POINTER(C.c_int)
    _chk(lib().orc_nms_topk(ph, B, Cat, H, W, K, int(heat_is_logit), score.ctypes.data_as(Pf), ind.ctypes.data_as(Pi),
                            cls.ctypes.data_as(Pi), ys.ctypes.data_as(Pf), xs.ctypes.data_as(Pf)), "orc_nms_topk")
    return score, ind, cls, ys, xs


def bbox_decode(hm, wh, reg, K=100, wh_scale=1.0):
    hm, ph = _f(hm); wh, pw = _f(wh); reg, pr = _f(reg)
    B, Cat, H, W = hm.shape
    bbox = np.empty((B, K, 5), np.float32); bbr = np.empty((B, K, 5), np.float32); keep = np.empty((B * K,), np.uint8)
    Pf = C.POINTER(C.c_float)
    _chk(lib().orc_bbox_decode(ph, pw, pr, B, Cat, H, W, K, C.c_float(wh_scale), bbox.ctypes.data_as(Pf),
                               bbr.ctypes.data_as(Pf), keep.ctypes.data_as(C.POINTER(C.c_uint8))), "orc_bbox_decode")
    return bbox, bbr, keep.astype(bool)


def ddd_decode(heat, kept, dim, orien, wh, reg, grid, K=40):
    heat, ph = _f(heat); kept, pk = _f(kept); dim, pd = _f(dim); orien, po = _f(orien); wh, pw = _f(wh); reg, pr = _f(reg)
    B, Cat, H, W = heat.shape
    assert kept.shape[1] == 6 * grid
    det = np.empty((B, K, 6), np.float32); detr = np.empty((B, K, 6), np.float32); info = np.empty((B, K, 9), np.float32)
    Pf = C.POINTER(C.c_float)
    _chk(lib().orc_ddd_decode(ph, pk, pd, po, pw, pr, B, Cat, H, W, grid, K, det.ctypes.data_as(Pf),
                              detr.ctypes.data_as(Pf), info.ctypes.data_as(Pf)), "orc_ddd_decode")
    return det, detr, info


def concat_volume(L, R, D):
    L, pL = _f(L); R, pR = _f(R)
    B, Cc, H, W = L.shape
    vol = np.empty((B, 2 * Cc, D, H, W), np.float32)
    _chk(lib().orc_concat_volume(pL, pR, vol.ctypes.data_as(C.POINTER(C.c_float)), B, Cc, H, W, D), "orc_concat_volume")
    return vol


def gwc_volume(L, R, D, G):
    L, pL = _f(L); R, pR = _f(R)
    B, Cc, H, W = L.shape
    vol = np.empty((B, G, D, H, W), np.float32)
    _chk(lib().orc_gwc_volume(pL, pR, vol.ctypes.data_as(C.POINTER(C.c_float)), B, Cc, H, W, D, G), "orc_gwc_volume")
    return vol


def conv3d_bn_relu(x, w, scale=None, shift=None, relu=False, residual=None):
    """x [N,Cin,D,H,W], w [Cout,Cin,3,3,3] -> [N,Cout,D,H,W]; folded eval-mode BatchNorm, ReLU, residual."""
    x, px = _f(x); w, pw = _f(w)
    N, Cin, D, H, W = x.shape
    Cout = w.shape[0]
    Pf = C.POINTER(C.c_float)
    ps = pt = pr = None
    if scale is not None:
        scale, ps = _f(scale); shift, pt = _f(shift)
    if residual is not None:
        residual, pr = _f(residual)
    y = np.empty((N, Cout, D, H, W), np.float32)
    _chk(lib().orc_conv3d_bn_relu(px, pw, ps, pt, pr, y.ctypes.data_as(Pf), N, Cin, D, H, W, Cout, 1 if relu else 0),
         "orc_conv3d_bn_relu")
    return y


def maxpool_hw2(x):
    x, px = _f(x)
    N, Cc, D, H, W = x.shape
    y = np.empty((N, Cc, D, H // 2, W // 2), np.float32)
    _chk(lib().orc_maxpool_hw2(px, y.ctypes.data_as(C.POINTER(C.c_float)), N * Cc, D, H, W), "orc_maxpool_hw2")
    return y


# ---- F2: dense photometric alignment (dense_align/dense_align.py) ----
def da_prep_u8(img, mean, std):
    """uint8 HWC image -> normalised, 2x bilinearly up-sampled planar [3, 2H, 2W] (align_parallel :251-266)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W, _ = img.shape
    mean, pm = _f(mean); std, ps = _f(std)
    out = np.empty((3, 2 * H, 2 * W), np.float32)
    _chk(lib().orc_da_prep_u8(img.ctypes.data_as(C.POINTER(C.c_ubyte)), pm, ps, out.ctypes.data_as(C.POINTER(C.c_float)),
                              H, W), "orc_da_prep_u8")
    return out


def da_sample(box_left, borders, poses, f, cx, cy, f_h, f_w, cap=8192):
    """sample() (:14-70): returns (uvz [rois, max, 3], weight [rois, max], count [rois]) trimmed to the longest RoI."""
    box_left, pb = _f(box_left); borders, pd = _f(borders); poses, pp = _f(poses)
    rois = box_left.shape[0]
    uvz = np.empty((rois, cap, 3), np.float32); wgt = np.empty((rois, cap), np.float32)
    cnt = np.empty((rois,), np.int32)
    P = C.POINTER(C.c_float)
    _chk(lib().orc_da_sample(pb, pd, pp, rois, C.c_float(f), C.c_float(cx), C.c_float(cy), f_h, f_w, cap,
                             uvz.ctypes.data_as(P), wgt.ctypes.data_as(P), cnt.ctypes.data_as(C.POINTER(C.c_int))),
         "orc_da_sample")
    assert cnt.max(initial=0) <= cap
    m = int(cnt.max(initial=0))
    return np.ascontiguousarray(uvz[:, :m]), np.ascontiguousarray(wgt[:, :m]), cnt


def da_enum(imL, imR, uvz, weight, depth_enum, fb, align_corners=False):
    """enumeration_depth (:175-237): returns (err_sum [iters, rois], best_depth [rois], best_idx [rois])."""
    imL, pl = _f(imL); imR, pr = _f(imR); uvz, pu = _f(uvz); weight, pw = _f(weight); depth_enum, pd = _f(depth_enum)
    H, W = imL.shape[-2:]
    rois, pixels = weight.shape
    iters = depth_enum.shape[0]
    err = np.empty((iters, rois), np.float32); best = np.empty((rois,), np.float32); idx = np.empty((rois,), np.int32)
    P = C.POINTER(C.c_float)
    _chk(lib().orc_da_enum(pl, pr, pu, pw, pd, C.c_float(fb), rois, pixels, iters, H, W, 1 if align_corners else 0,
                           err.ctypes.data_as(P), best.ctypes.data_as(P), idx.ctypes.data_as(C.POINTER(C.c_int))),
         "orc_da_enum")
    return err, best, idx


# ---- F3: stereo_network_new voxel volume (models/networks/stereo_network_new.py:160-283, 409-449) ----
def voxel_coords(left, right, p2, p3, fb, trans, trans_inv, depth_bins, input_h=384, input_w=1280):
    left, pl = _f(left); right, pr = _f(right); p2, pp2 = _f(p2); p3, pp3 = _f(p3); fb, pf = _f(fb)
    trans, pt = _f(trans); trans_inv, pti = _f(trans_inv); depth_bins, pd = _f(depth_bins)
    N, B, D = left.shape[0], fb.shape[0], depth_bins.shape[1]
    P = C.POINTER(C.c_float)
    norm3 = np.empty((N, 10, 10, 10, 3), np.float32); valid3 = np.empty((N, 10, 10, 10), np.float32)
    normL = np.empty((N, 10, 10, 10, 2), np.float32); validL = np.empty((N, 10, 10, 10), np.float32)
    normR = np.empty((N, 10, 10, 10, 2), np.float32); validR = np.empty((N, 10, 10, 10), np.float32)
    dori = np.empty((N,), np.float32)
    _chk(lib().orc_voxel_coords(pl, pr, pp2, pp3, pf, pt, pti, pd, N, B, D, input_h, input_w, norm3.ctypes.data_as(P),
                                valid3.ctypes.data_as(P), normL.ctypes.data_as(P), validL.ctypes.data_as(P),
                                normR.ctypes.data_as(P), validR.ctypes.data_as(P), dori.ctypes.data_as(P)), "orc_voxel_coords")
    return norm3, valid3, normL, validL, normR, validR, dori


def voxel_volume(featL, featR, left, right, p2, p3, fb, trans, trans_inv, input_h=384, input_w=1280, align_corners=False):
    featL, pfl = _f(featL); featR, pfr = _f(featR)
    left, pl = _f(left); right, pr = _f(right); p2, pp2 = _f(p2); p3, pp3 = _f(p3); fb, pf = _f(fb)
    trans, pt = _f(trans); trans_inv, pti = _f(trans_inv)
    N = left.shape[0]
    B, Cc, H, W = featL.shape
    P = C.POINTER(C.c_float)
    voxel = np.empty((N, 3 * Cc, 10, 10, 10), np.float32); dori = np.empty((N,), np.float32)
    _chk(lib().orc_voxel_volume(pfl, pfr, pl, pr, pp2, pp3, pf, pt, pti, N, B, Cc, H, W, input_h, input_w,
                                1 if align_corners else 0, voxel.ctypes.data_as(P), dori.ctypes.data_as(P)), "orc_voxel_volume")
    return voxel, dori
