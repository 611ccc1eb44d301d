"""Synthetic KITTI-shaped inputs and a non-degenerate random initialisation.

The reference constructors' default init makes every DCN offset 0 / mask 0.5 (``DCN.init_offset``,
dcn_v2.py:112-116) and lets activations vanish to ~1e-6 through the eval-mode BatchNorms, so every box is
degenerate.  ``realistic_init`` keeps the architecture and parameter shapes but draws weights that keep
activations O(1), offsets ~N(0, 1 px) and boxes of plausible size, so the gather / RoI / top-K kernels see
realistic access patterns.  It is deterministic (CPU generator) and shared by tests, bench and the golden
vector generator.
"""
import math

import torch
from torch import nn

HEADS = {'hm': 3, 'dim': 3, 'orien': 2, 'kept_type': 168, 'wh': 3, 'reg': 3}   # opts.py:304-311, grid=28
KITTI_FB = 384.38   # f * baseline = 721.54 * 0.5327


def make_batch(B=1, H=384, W=1280, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    p2 = torch.tensor([[721.54, 0, 609.56, 44.86], [0, 721.54, 172.85, 0.216], [0, 0, 1, 0.00275]])
    p3 = p2.clone()
    p3[0, 3] = -339.52
    batch = {
        'input': torch.randn(B, 3, H, W, generator=g),
        'input_right': torch.randn(B, 3, H, W, generator=g),
        'fb': torch.full((B,), KITTI_FB),
        'p2': p2.repeat(B, 1, 1), 'p3': p3.repeat(B, 1, 1),
        'trans': torch.tensor([[.25, 0, 0], [0, .25, 0]]).repeat(B, 1, 1),
        'trans_inv': torch.tensor([[4., 0, 0], [0, 4., 0]]).repeat(B, 1, 1),
    }
    return {k: v.to(device) for k, v in batch.items()}


def make_boxes(B, n_per_image, seed=0, W4=320, H4=96):
    """Left/right boxes as SURVEY.md config #2: x1~U(20,270), w~U(8,58), y1~U(20,70), h~U(6,31),
    right box = left shifted by U(2,12) px.  Returns (left [N,5], right [N,5], torch.Size([B,n,5]))."""
    g = torch.Generator().manual_seed(seed)
    N = B * n_per_image
    x1 = 20 + 250 * torch.rand(N, generator=g)
    w = 8 + 50 * torch.rand(N, generator=g)
    y1 = 20 + 50 * torch.rand(N, generator=g)
    h = 6 + 25 * torch.rand(N, generator=g)
    sx, sy = W4 / 320.0, H4 / 96.0
    x1, w, y1, h = x1 * sx, w * sx, y1 * sy, h * sy
    shift = (2 + 10 * torch.rand(N, generator=g)) * sx
    b = torch.arange(B).repeat_interleave(n_per_image).float()
    left = torch.stack([b, x1, y1, x1 + w, y1 + h], 1)
    right = torch.stack([b, x1 - shift, y1, x1 + w - shift, y1 + h], 1)
    return left, right, torch.Size([B, n_per_image, 5])


def make_targets(B, n_obj=8, max_objs=50, seed=0, W4=320, H4=96, grid=28):
    """Synthetic ground truth of one training batch in the layout of stereoDataset.__getitem__ (modules/stereoDataset.py:272-285;
    SURVEY.md config #5: 8 objects per pair, boxes as config #2): 'hm' [B,3,H4,W4] with a radius-2 gaussian per object,
    'ind' / 'ind_float' [B,M] flat peak index, 'wh' [B,M,3] = (w_left, w_right, h), 'reg' [B,M,3] = (dx_left, dx_right, dy),
    'dim', 'orien', 'depth', 'kept' [B,M,6], 'rot_mask' [B,M]; rows past n_obj are zero."""
    g = torch.Generator().manual_seed(seed)
    left, right, _ = make_boxes(B, n_obj, seed=seed, W4=W4, H4=H4)
    left, right = left.view(B, n_obj, 5), right.view(B, n_obj, 5)
    M = max_objs
    t = {'hm': torch.zeros(B, 3, H4, W4), 'ind': torch.zeros(B, M, dtype=torch.long), 'ind_float': torch.zeros(B, M),
         'wh': torch.zeros(B, M, 3), 'reg': torch.zeros(B, M, 3), 'dim': torch.zeros(B, M, 3), 'orien': torch.zeros(B, M, 2),
         'depth': torch.zeros(B, M, 1), 'kept': torch.zeros(B, M, 6), 'rot_mask': torch.zeros(B, M)}
    cxl, cxr = 0.5 * (left[..., 1] + left[..., 3]), 0.5 * (right[..., 1] + right[..., 3])
    cy = 0.5 * (left[..., 2] + left[..., 4])
    xi, yi = cxl.floor().clamp(0, W4 - 1), cy.floor().clamp(0, H4 - 1)
    ind = (yi * W4 + xi).long()
    t['ind'][:, :n_obj] = ind
    t['ind_float'][:, :n_obj] = ind.float()
    t['wh'][:, :n_obj] = torch.stack((left[..., 3] - left[..., 1], right[..., 3] - right[..., 1], left[..., 4] - left[..., 2]), 2)
    t['reg'][:, :n_obj] = torch.stack((cxl - xi, cxr - xi, cy - yi), 2)
    t['dim'][:, :n_obj] = torch.rand(B, n_obj, 3, generator=g) * 2 + 1
    t['orien'][:, :n_obj] = torch.randn(B, n_obj, 2, generator=g)
    t['depth'][:, :n_obj] = torch.rand(B, n_obj, 1, generator=g) * 55 + 5
    t['kept'][:, :n_obj] = torch.rand(B, n_obj, 6, generator=g) * t['wh'][:, :n_obj, :1]
    t['rot_mask'][:, :n_obj] = 1
    cls = torch.randint(0, 3, (B, n_obj), generator=g)
    yy, xx = torch.meshgrid(torch.arange(-2, 3), torch.arange(-2, 3), indexing="ij")
    gauss = torch.exp(-(xx * xx + yy * yy).float() / (2 * (5 / 6.0) ** 2))
    for b in range(B):
        for o in range(n_obj):
            x, y, c = int(xi[b, o]), int(yi[b, o]), int(cls[b, o])
            x0, x1, y0, y1 = max(x - 2, 0), min(x + 3, W4), max(y - 2, 0), min(y + 3, H4)
            patch = t['hm'][b, c, y0:y1, x0:x1]
            torch.maximum(patch, gauss[y0 - y + 2:y1 - y + 2, x0 - x + 2:x1 - x + 2], out=patch)
    return t


def _calibrate_batchnorm(model, g):
    """Data-dependent BatchNorm statistics so that eval-mode activations stay O(1) (pure torch, CPU).

    One train-mode pass with momentum 1 over a small seeded input sets running_mean / running_var to the
    batch statistics.  The DCN layers are replaced by their zero-offset surrogate 0.5*conv2d(x, W) + b for this
    pass only (exact for the reference's default init) so no deformable kernel is needed.
    """
    import torch.nn.functional as F
    from ..dcn_v2 import DCN
    from ..networks.feature_extraction_dla34 import DepthwiseUp

    bns = [m for m in model.modules() if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d))]
    saved = [(m.momentum, m.training) for m in bns]
    orig_forward = DCN.forward
    orig_up = DepthwiseUp.forward

    def surrogate(self, input, bn=None, relu=False):
        return 0.5 * F.conv2d(input, self.weight, None, self.stride, self.padding, self.dilation) + \
            self.bias.view(1, -1, 1, 1)

    try:
        DCN.forward = surrogate
        DepthwiseUp.forward = nn.ConvTranspose2d.forward      # plain ATen op for this CPU-side pass
        for m in bns:
            m.momentum = 1.0
            m.train()
        with torch.no_grad():
            x = torch.randn(2, 3, 192, 640, generator=g)
            f = model.feature_extraction(x)
            model.feaRuduce(f)
            est = model.depth_estimator
            cost = 0.5 * torch.randn(4, 3 * est.reduced_channel, 16, 16, 16, generator=g)
            est.aggregate(cost)
    finally:
        DCN.forward = orig_forward
        DepthwiseUp.forward = orig_up
        for m, (mom, tr) in zip(bns, saved):
            m.momentum = mom
            m.train(tr)


def realistic_init(model, seed=0):
    g = torch.Generator().manual_seed(seed)

    def normal_(t, std):
        with torch.no_grad():
            t.copy_(torch.randn(t.shape, generator=g) * std)

    dcn_like = [m for m in model.modules() if hasattr(m, "conv_offset_mask")]
    skip = set()
    for m in dcn_like:
        fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
        normal_(m.weight, 2.0 * math.sqrt(2.0 / fan_in))           # x2: the mask averages 0.5
        normal_(m.bias, 0.05)
        normal_(m.conv_offset_mask.weight, 1.0 / math.sqrt(fan_in))
        normal_(m.conv_offset_mask.bias, 0.1)
        skip.add(m.conv_offset_mask)
    for m in model.modules():
        if isinstance(m, (nn.Conv2d, nn.Conv3d)) and m not in skip:
            fan_in = m.in_channels // m.groups
            for k in m.kernel_size:
                fan_in *= k
            normal_(m.weight, math.sqrt(2.0 / fan_in))
        elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
            with torch.no_grad():
                m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
    _calibrate_batchnorm(model, g)
    # heads: plausible box sizes / offsets / stereo shift, distinct heat-map peaks
    with torch.no_grad():
        if hasattr(model, "wh"):
            normal_(model.wh[-1].weight, 0.2)
            model.wh[-1].bias.copy_(torch.tensor([24.0, 24.0, 14.0]))
        if hasattr(model, "reg"):
            normal_(model.reg[-1].weight, 0.01)
            model.reg[-1].bias.copy_(torch.tensor([0.5, -5.0, 0.5]))
        if hasattr(model, "hm"):
            normal_(model.hm[-1].weight, 0.05)
            model.hm[-1].bias.fill_(-2.19)
        for name in ("dim", "orien", "kept_type"):
            if hasattr(model, name):
                normal_(getattr(model, name)[-1].weight, 0.05)
    return model
