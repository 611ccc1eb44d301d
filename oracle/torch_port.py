"""TEST INFRASTRUCTURE ONLY -- PyTorch (CPU-capable) restatement of the reference's op SEQUENCE.

Where ``side_oracle.c`` restates the arithmetic of single operators, this file restates how the
reference strings library ops together on the hot path -- torchvision ``deform_conv2d`` (BASELINE.json
names it as the CPU stand-in for the legacy DCNv2 extension), a Python loop of 2*D torchvision
``roi_align`` calls with slice assignment, ``torch.topk``-based decode, softmax + D-step loop -- so that

  * tests can check the product (CUDA) against the reference's behaviour end to end, including
    gradients through ``torch.autograd``;
  * ``bench.py`` can time "the reference's CPU path" on the GPU box, where ``/root/reference`` does not
    exist (``cpu_baseline.kind = "port"``).

``reference_ops()`` temporarily swaps the product's operator entry points (``side_b200.ops``) for these
restatements, which turns the product's ``nn.Module`` tree (pure torch structure, identical state dict)
into a runnable port of ``stereo_network_old``.  It is only ever entered from ``tests/`` and from
``bench.py``'s CPU legs; the product itself never imports this package.

Pinned against the real reference in the build container by ``tests/test_reference_live.py``
and the committed golden vectors (``oracle/gen_golden.py``).
"""
import contextlib

import torch
import torch.nn.functional as F
import torchvision.ops as tvo


# ---------------------------------------------------------------------------------------------
# DCN  (reference: DCNv2/dcn_v2.py:16-51 -> _ext; CPU stand-in torchvision.ops.deform_conv2d)
# ---------------------------------------------------------------------------------------------
def dcn_v2_conv(input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups):
    return tvo.deform_conv2d(input, offset, weight, bias, stride=stride, padding=padding, dilation=dilation, mask=mask)


def dcn_module_forward(input, om, weight, bias, stride, padding, dilation):
    """DCN.forward, dcn_v2.py:118-128."""
    o1, o2, mask = torch.chunk(om, 3, dim=1)
    offset = torch.cat((o1, o2), dim=1)
    mask = torch.sigmoid(mask)
    return dcn_v2_conv(input, offset, mask, weight, bias, stride, padding, dilation, 1)


def dcn_fused_infer(input, om, weight, bias, stride, padding, dilation, scale=None, shift=None, relu=False):
    y = dcn_module_forward(input, om, weight, bias, stride, padding, dilation)
    if scale is not None:   # eval-mode BatchNorm written as an affine map
        y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    return F.relu(y) if relu else y


# ---------------------------------------------------------------------------------------------
# proposals / instance volume / gate / soft-argmin  (stereo_network_old.py)
# ---------------------------------------------------------------------------------------------
def proposal_shift(left, right, fb, D, x_clamp):
    """get_proposal_shift, :34-133 (boxes already grouped by image)."""
    rate = torch.tensor([float(i) / (D - 1) for i in range(D)], dtype=torch.float32, device=left.device)
    b = left[:, 0].long()
    fbn = fb[b]
    xmin = torch.min(left[:, 1], right[:, 1])
    ymin = torch.min(left[:, 2], right[:, 2])
    xmax = torch.max(left[:, 3], right[:, 3])
    ymax = torch.max(left[:, 4], right[:, 4])
    dmin = (fbn / ((xmax - xmin) * 0.9 * 4)).view(-1, 1)
    dmin = torch.clamp(dmin, min=1.0, max=87.0)
    depth_bin = 87.0 - (87.0 - dmin) * rate
    disp = fbn.view(-1, 1) / depth_bin / 8
    pl, pr = [], []
    for i in range(D):
        pl.append(torch.stack((left[:, 0], torch.clamp(xmin + disp[:, i], max=x_clamp), ymin,
                               torch.clamp(xmax + disp[:, i], max=x_clamp), ymax), dim=1))
        pr.append(torch.stack((left[:, 0], torch.clamp(xmin - disp[:, i], min=0.), ymin,
                               torch.clamp(xmax - disp[:, i], min=0.), ymax), dim=1))
    return torch.stack(pl, 0), torch.stack(pr, 0), depth_bin


def xcross_gate(cost, C):
    """cost_volume.forward :197-203 with num_channels generalised to C."""
    l, r = cost[:, :C], cost[:, C:2 * C]
    ln = torch.sqrt(torch.sum(l * l, (1, 3, 4)))
    rn = torch.sqrt(torch.sum(r * r, (1, 3, 4)))
    xc = torch.sum(l * r, (1, 3, 4)) / torch.clamp(ln * rn, min=0.01)
    return cost * xc.unsqueeze(1).unsqueeze(3).unsqueeze(4)


def inst_costvol(featL, featR, left, right, fb, D, P, x_clamp, gate=False, valid=None):
    """:366-376: D x (RoIAlign left, RoIAlign right) + slice assignment of [L, R, L-R]."""
    pl, pr, depth_bin = proposal_shift(left, right, fb, D, x_clamp)
    N, C = left.shape[0], featL.shape[1]
    cost = torch.zeros((N, 3 * C, D, P, P), dtype=featL.dtype, device=featL.device)
    for i in range(D):
        rl = tvo.roi_align(featL, pl[i], (P, P), spatial_scale=1, sampling_ratio=2)
        rr = tvo.roi_align(featR, pr[i], (P, P), spatial_scale=1, sampling_ratio=2)
        cost[:, :C, i] = rl
        cost[:, C:2 * C, i] = rr
        cost[:, 2 * C:, i] = rl - rr
    if gate:
        cost = xcross_gate(cost, C)
    if valid is not None:
        v = valid.bool()
        cost = cost * v.view(-1, 1, 1, 1, 1).to(cost.dtype)
        depth_bin = depth_bin * v.view(-1, 1).to(cost.dtype)
    return cost, depth_bin


def softargmin(logits, depth_bin):
    """:228-236: AvgPool2d(4) -> softmax over D -> sequential sum of p_i * depth_bin_i."""
    N, D = logits.shape[:2]
    pooled = F.avg_pool2d(logits, logits.shape[-1], logits.shape[-1])
    pred = F.softmax(pooled.reshape(N, D), dim=1)
    disp = torch.zeros(N, dtype=logits.dtype, device=logits.device)
    for i in range(D):
        disp = disp + pred[:, i] * depth_bin[:, i]
    return disp


# ---------------------------------------------------------------------------------------------
# decode  (models/decode.py, models/utils.py)
# ---------------------------------------------------------------------------------------------
def _nms(heat):
    hmax = F.max_pool2d(heat, (3, 3), stride=1, padding=1)
    return heat * (hmax == heat).float()


def _gather(feat, ind):
    """_transpose_and_gather_feat, utils.py:21-26."""
    B, Cc = feat.shape[:2]
    f = feat.permute(0, 2, 3, 1).contiguous().view(B, -1, Cc)
    return f.gather(1, ind.unsqueeze(2).expand(B, ind.shape[1], Cc))


def _topk(scores, K):
    """decode.py:17-33."""
    B, cat, H, W = scores.shape
    s1, i1 = torch.topk(scores.view(B, cat, -1), K)
    i1 = i1 % (H * W)
    ys1 = torch.div(i1, W, rounding_mode='floor').float()
    xs1 = (i1 % W).float()
    s2, i2 = torch.topk(s1.view(B, -1), K)
    cls = torch.div(i2, K, rounding_mode='floor').int()
    ind = i1.view(B, -1).gather(1, i2)
    return s2, ind, cls, ys1.view(B, -1).gather(1, i2), xs1.view(B, -1).gather(1, i2)


def bbox_decode_raw(heat, wh, reg, K=100, wh_scale=1.0, heat_is_logit=True):
    """bbox_decode, decode.py:91-126, without the final compaction."""
    if heat_is_logit:
        heat = torch.sigmoid(heat)
    score, ind, cls, ys, xs = _topk(_nms(heat), K)
    B = heat.shape[0]
    r = _gather(reg, ind)
    w = _gather(wh * wh_scale, ind)
    cx, cxr, cy = xs + r[:, :, 0], xs + r[:, :, 1], ys + r[:, :, 2]
    bidx = torch.arange(B, dtype=torch.float32, device=heat.device).view(B, 1).expand(B, K)
    bbox = torch.stack([bidx, cx - 0.5 * w[:, :, 0], cy - 0.5 * w[:, :, 2], cx + 0.5 * w[:, :, 0], cy + 0.5 * w[:, :, 2]], 2)
    bbr = torch.stack([bidx, cxr - 0.5 * w[:, :, 1], cy - 0.5 * w[:, :, 2], cxr + 0.5 * w[:, :, 1], cy + 0.5 * w[:, :, 2]], 2)
    keep = torch.sum(bbox.view(-1, 5)[:, 1:5], dim=1) > 0
    kb = keep.view(B, K)
    slot = torch.cumsum(kb.int(), 1) - kb.int()
    return dict(bbox=bbox, bbox_right=bbr, keep=keep.to(torch.uint8), slot=slot.int(), count=kb.sum(1).int(), score=score,
                ind=ind.int(), cls=cls)


def ddd_decode_raw(heat, kept, dim, orien, wh, reg, grid_size, K=40, heat_is_logit=False):
    """ddd_decode, decode.py:35-89 (kept_type = floor(argmax/grid))."""
    if heat_is_logit:
        heat = torch.sigmoid(heat)
    score, ind, cls, ys, xs = _topk(_nms(heat), K)
    r = _gather(reg, ind)
    w = _gather(wh, ind)
    g = grid_size
    a0 = _gather(kept[:, :4 * g], ind).argmax(2)
    a1 = _gather(kept[:, 4 * g:5 * g], ind).argmax(2)
    a2 = _gather(kept[:, 5 * g:], ind).argmax(2)
    clsf = cls.float().unsqueeze(2)
    det = torch.cat([(xs + r[:, :, 0]).unsqueeze(2), (ys + r[:, :, 2]).unsqueeze(2), w[:, :, [0, 2]], score.unsqueeze(2), clsf], 2)
    detr = torch.cat([(xs + r[:, :, 1]).unsqueeze(2), (ys + r[:, :, 2]).unsqueeze(2), w[:, :, [1, 2]], score.unsqueeze(2), clsf], 2)
    info = torch.cat([_gather(dim, ind), _gather(orien, ind), a1.float().unsqueeze(2), a2.float().unsqueeze(2),
                      (a0 % g).float().unsqueeze(2), torch.div(a0, g, rounding_mode='floor').float().unsqueeze(2)], 2)
    return det, detr, info


# ---------------------------------------------------------------------------------------------
# full-image volumes (no reference call site: PARITY UNPINNED, semantics in include/side_b200.h)
# ---------------------------------------------------------------------------------------------
def concat_volume(L, R, D):
    B, C, H, W = L.shape
    vol = L.new_zeros((B, 2 * C, D, H, W))
    for d in range(D):
        vol[:, :C, d, :, d:] = L[:, :, :, d:]
        vol[:, C:, d, :, d:] = R[:, :, :, :W - d]
    return vol


def gwc_volume(L, R, D, G):
    B, C, H, W = L.shape
    vol = L.new_zeros((B, G, D, H, W))
    for d in range(D):
        prod = (L[:, :, :, d:] * R[:, :, :, :W - d]).view(B, G, C // G, H, W - d).mean(2)
        vol[:, :, d, :, d:] = prod
    return vol


def dw_deconv(x, w, stride, pad):
    """IDAUp.up_k, feature_extraction_dla34.py:370-373: the ATen/cuDNN call the reference makes."""
    return F.conv_transpose2d(x, w, None, stride=stride, padding=pad, groups=x.shape[1])


# ---- F3: stereo_network_new voxel volume (models/networks/stereo_network_new.py:160-283, 409-449) ----
def voxel_volume(featL, featR, left, right, p2, p3, fb, trans, trans_inv, input_h=384, input_w=1280):
    """The reference's op sequence on torch tensors (any device): get_voxel's projection of the 10^3 metric grid into both views,
    F.grid_sample per RoI, validity masks, cat(L - R, L, R).  Returns (voxel [N, 3C, 10, 10, 10], depth_ori [N])."""
    F = torch.nn.functional
    N = left.shape[0]
    bi = left[:, 0].long()
    ti, P2, P3, tr = trans_inv.float()[bi], p2.float()[bi], p3.float()[bi], trans.float()[bi]
    ox = lambda x, y: x * ti[:, 0, 0] + y * ti[:, 0, 1] + ti[:, 0, 2]
    oy = lambda x, y: x * ti[:, 1, 0] + y * ti[:, 1, 1] + ti[:, 1, 2]
    cx = (ox(left[:, 1], left[:, 2]) + ox(left[:, 3], left[:, 4])) / 2
    cy = (oy(left[:, 1], left[:, 2]) + oy(left[:, 3], left[:, 4])) / 2
    cxr = (ox(right[:, 1], right[:, 2]) + ox(right[:, 3], right[:, 4])) / 2
    depth = fb.float().reshape(-1)[bi] / (cx - cxr)
    z = depth - P2[:, 2, 3]
    x = (cx * depth - P2[:, 0, 3] - P2[:, 0, 2] * z) / P2[:, 0, 0]
    y = (cy * depth - P2[:, 1, 3] - P2[:, 1, 2] * z) / P2[:, 1, 1]
    stride = 0.5
    dev = left.device
    xs = (torch.arange(-2.5, 2.5, stride, device=dev) + stride / 2).view(1, 10, 1, 1) + x.view(N, 1, 1, 1)
    ys = (torch.arange(-2.5, 2.5, stride, device=dev) + stride / 2).view(1, 1, 10, 1) + y.view(N, 1, 1, 1)
    zs = (torch.arange(-5., 5., 1., device=dev) + 0.5).view(1, 1, 1, 10) + z.view(N, 1, 1, 1)
    X, Y, Z = torch.broadcast_tensors(xs, ys, zs)

    def view(P):
        e = lambda r, c: P[:, r, c].view(N, 1, 1, 1)
        a = X * e(0, 0) + Y * e(0, 1) + Z * e(0, 2) + e(0, 3)
        b = X * e(1, 0) + Y * e(1, 1) + Z * e(1, 2) + e(1, 3)
        w = X * e(2, 0) + Y * e(2, 1) + Z * e(2, 2) + e(2, 3)
        u, v, one = a / w, b / w, w / w
        t = lambda r, c: tr[:, r, c].view(N, 1, 1, 1)
        uf = u * t(0, 0) + v * t(0, 1) + one * t(0, 2)
        vf = u * t(1, 0) + v * t(1, 1) + one * t(1, 2)
        g = torch.stack([uf / (input_w / 4 - 1.) * 2. - 1., vf / (input_h / 4 - 1.) * 2. - 1.], -1)       # [N,10,10,10,2]
        ok = ((g[..., 0] >= -1.) & (g[..., 0] <= 1.) & (g[..., 1] >= -1.) & (g[..., 1] <= 1.)).float()
        return g * ok.unsqueeze(4), ok
    gl, okl = view(P2)
    gr, okr = view(P3)
    L = torch.cat([F.grid_sample(featL[int(bi[n]):int(bi[n]) + 1], gl[n].reshape(1, 1, -1, 2), align_corners=False)
                   for n in range(N)], 0).reshape(N, -1, 10, 10, 10) * okl.unsqueeze(1)
    R = torch.cat([F.grid_sample(featR[int(bi[n]):int(bi[n]) + 1], gr[n].reshape(1, 1, -1, 2), align_corners=False)
                   for n in range(N)], 0).reshape(N, -1, 10, 10, 10) * okr.unsqueeze(1)
    return torch.cat([L - R, L, R], 1), depth


# ---------------------------------------------------------------------------------------------
@contextlib.contextmanager
def reference_ops():
    """Swaps side_b200's operator entry points for the restatements above (tests / CPU baseline only)."""
    import side_b200.ops as ops
    import side_b200.dcn_v2 as dcn_mod

    def fused(input, om, weight, bias, stride, padding, dilation):
        return dcn_module_forward(input, om, weight, bias, stride, padding, dilation)

    repl = dict(dcn_v2_conv=dcn_v2_conv, dcn_fused=fused, dcn_fused_infer=dcn_fused_infer, proposal_shift=proposal_shift,
                inst_costvol=inst_costvol, xcross_gate=xcross_gate, softargmin=softargmin,
                bbox_decode_raw=bbox_decode_raw, ddd_decode_raw=ddd_decode_raw, concat_volume=concat_volume,
                gwc_volume=gwc_volume, dw_deconv=dw_deconv, voxel_volume=voxel_volume)
    saved = {k: getattr(ops, k) for k in repl}
    saved_dcn = dcn_mod.dcn_v2_conv
    try:
        for k, v in repl.items():
            setattr(ops, k, v)
        dcn_mod.dcn_v2_conv = dcn_v2_conv
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)
        dcn_mod.dcn_v2_conv = saved_dcn


# ---- F4: detector input preparation (modules/stereoDetector.py:45-82) ----
def warp_affine_u8(img, inv_map, out_hw):
    """TEST INFRASTRUCTURE.  numpy restatement of cv2.warpAffine(img, M, (w, h), flags=INTER_LINEAR) for 8-bit images with
    the default constant border 0, from OpenCV's published algorithm (imgwarp.cpp: WarpAffineInvoker + remapBilinear);
    ``inv_map`` is the destination -> source map (what warpAffine computes from M).  cv2 is not installed in the build
    image: PARITY UNPINNED against the real library."""
    import numpy as np
    img = np.asarray(img)
    H, W = img.shape[:2]
    h, w = out_hw
    m = np.asarray(inv_map, np.float64).reshape(-1)
    rnd = lambda v: np.clip(np.rint(v), -2147483648.0, 2147483647.0).astype(np.int64)
    xs, ys = np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64)
    X0 = rnd((m[1] * ys + m[2]) * 1024.0) + 16
    Y0 = rnd((m[4] * ys + m[5]) * 1024.0) + 16
    X = (X0[:, None] + rnd(m[0] * xs * 1024.0)[None, :]) >> 5
    Y = (Y0[:, None] + rnd(m[3] * xs * 1024.0)[None, :]) >> 5
    sx, sy, ax, ay = X >> 5, Y >> 5, X & 31, Y & 31
    out = np.zeros((h, w, img.shape[2]), np.int64)
    for dy, wy in ((0, 32 - ay), (1, ay)):
        for dx, wx in ((0, 32 - ax), (1, ax)):
            yy, xx = sy + dy, sx + dx
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            pix = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64) * ok[..., None]
            out += pix * (wy * wx * 32)[..., None]
    return np.clip((out + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def pre_process_ref(image, trans_input, out_hw, mean, std):
    """The host pipeline of stereoDetector.pre_process on top of ``warp_affine_u8``: float32 1 x 3 x h x w."""
    import numpy as np
    from side_b200.preprocess import invert_affine
    inp = warp_affine_u8(image, invert_affine(trans_input), out_hw)
    inp = (inp.astype(np.float32) / 255.)
    inp = (inp - np.array(mean, dtype=np.float32).reshape(1, 1, 3)) / np.array(std, dtype=np.float32).reshape(1, 1, 3)
    return inp.transpose(2, 0, 1)[np.newaxis, ...]
