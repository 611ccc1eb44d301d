"""Training-side host glue around the hot path (BASELINE config #5).

``ModelWithLoss.forward`` of the reference (modules/stereoTrainer.py:36-64) builds the ground-truth RoIs for the depth
branch on the CPU -- ``torch.zeros`` host tensors, a boolean-mask compaction and a ``.cuda()`` upload per step -- before it
calls the network, and ``StereoLoss`` (:66-144) combines the CenterNet focal loss, masked L1 regressions, the three keypoint
cross-entropies and the L1 depth loss.  Here the RoIs are built on the device in the fixed shape [B * max_objs] with a
validity mask (no host synchronisation: ``stereo_network.forward`` accepts the 4-tuple), and the loss keeps the reference's
terms, reductions and weights so the backward pass exercises the same operators with the same gradient magnitudes.
"""
import torch
import torch.nn.functional as F
from torch import nn


def gt_rois(batch, output_w, wh_scale=1.0):
    """stereoTrainer.py:41-63 on the device.  batch: 'ind_float' [B,M], 'wh' [B,M,3] (w_left, w_right, h), 'reg' [B,M,3]
    (dx_left, dx_right, dy).  -> (bbox [B*M,5], bbox_right [B*M,5], torch.Size([B,M,5]), keep uint8 [B*M]); rows are
    image-major, i.e. already grouped by image in ascending b as get_proposal_shift requires."""
    ind = batch['ind_float']
    wh, reg = batch['wh'], batch['reg']
    B, M = ind.shape
    xs = torch.remainder(ind, output_w)                         # ind_float % output_w
    ys = torch.div(ind, output_w, rounding_mode='floor')        # ind_float // output_w
    xs_right = xs + reg[:, :, 1]
    xs, ys = xs + reg[:, :, 0], ys + reg[:, :, 2]
    bi = torch.arange(B, device=ind.device, dtype=torch.float32).view(B, 1).expand(B, M)
    hw, hwr, hh = 0.5 * wh[:, :, 0] * wh_scale, 0.5 * wh[:, :, 1] * wh_scale, 0.5 * wh[:, :, 2] * wh_scale
    bbox = torch.stack((bi, xs - hw, ys - hh, xs + hw, ys + hh), 2)
    bbox_right = torch.stack((bi, xs_right - hwr, ys - hh, xs_right + hwr, ys + hh), 2)
    keep = (bbox[:, :, 1:5].sum(2) > 0).view(-1).to(torch.uint8)
    return bbox.view(-1, 5).contiguous(), bbox_right.view(-1, 5).contiguous(), torch.Size([B, M, 5]), keep


def _gather_feat(output, ind):
    """_transpose_and_gather_feat (models/utils.py:21-26) without the NHWC copy: [B,C,H,W], ind [B,M] -> [B,M,C]."""
    B, C = output.shape[:2]
    return output.flatten(2).gather(2, ind.unsqueeze(1).expand(B, C, ind.shape[1])).transpose(1, 2)


def focal_loss(pred, gt):
    """_neg_loss (models/losses.py:42-67): CornerNet focal loss on the sigmoided heat map."""
    pos = gt.eq(1).float()
    neg = gt.lt(1).float()
    neg_w = torch.pow(1 - gt, 4)
    pos_loss = (torch.log(pred) * torch.pow(1 - pred, 2) * pos).sum()
    neg_loss = (torch.log(1 - pred) * torch.pow(pred, 2) * neg_w * neg).sum()
    num_pos = pos.sum()
    # reference: `if num_pos == 0` on the host; branch-free here so the step never synchronises
    return torch.where(num_pos == 0, -neg_loss, -(pos_loss + neg_loss) / num_pos.clamp_min(1.0))


def reg_l1(output, mask, ind, target):
    """L1Loss (models/losses.py:177-185): mean over ALL B*M*C entries of |pred*mask - target*mask|."""
    pred = _gather_feat(output, ind)
    m = mask.unsqueeze(2).expand_as(pred).float()
    return F.l1_loss(pred * m, target * m, reduction='mean')


def cross_loss(output, ind, target):
    """CrossLoss (models/losses.py:187-198): cross entropy of the gathered logits, mean over all B*M rows (unmasked)."""
    pred = _gather_feat(output, ind)
    return F.cross_entropy(pred.reshape(-1, pred.shape[2]), target.reshape(-1).long(), reduction='mean')


def kept_label(kept, wh, grid):
    """StereoLoss.computeKeptLabel (stereoTrainer.py:75-93): [B,M,6] keypoint offsets -> [B,M,3] class indices."""
    width = (wh[:, :, 0] + 1).unsqueeze(2).expand(-1, -1, 6)
    target = torch.round(kept * grid / width)
    target = torch.where((target < 0) | (target > grid - 1), torch.full_like(target, -225.0), target)
    pos, typ = torch.max(target[:, :, :4], 2)
    target = torch.cat(((typ.float() * grid + pos).unsqueeze(2), target[:, :, 4:]), 2)
    return torch.clamp_min(target, 0).long()


class StereoLoss(nn.Module):
    """The reference's loss composition (stereoTrainer.py:66-144), default options: focal heat-map loss, fixed weights."""

    def __init__(self, grid=28, loss_weight=(1., 1., 1., 1., 1., 1., 1.), cost_volume=True):
        super().__init__()
        self.grid, self.w, self.cost_volume = grid, tuple(loss_weight), cost_volume

    def forward(self, outputs, batch):
        out = outputs[-1]
        g = self.grid
        depth_loss = F.l1_loss(out['depth'], batch['depth'], reduction='mean') if self.cost_volume else 0.0
        hm = torch.clamp(torch.sigmoid(out['hm']), min=1e-4, max=1 - 1e-4)          # models/utils.py:_sigmoid
        hm_loss = focal_loss(hm, batch['hm'])
        dim_loss = reg_l1(out['dim'], batch['rot_mask'], batch['ind'], batch['dim'])
        orien_loss = reg_l1(out['orien'], batch['rot_mask'], batch['ind'], batch['orien'])
        target = kept_label(batch['kept'], batch['wh'], g)
        kt = out['kept_type']
        kept_loss = (cross_loss(kt[:, :4 * g], batch['ind'], target[:, :, 0]) + cross_loss(kt[:, 4 * g:5 * g], batch['ind'], target[:, :, 1]) +
                     cross_loss(kt[:, 5 * g:], batch['ind'], target[:, :, 2])) / 3
        wh_loss = reg_l1(out['wh'], batch['rot_mask'], batch['ind'], batch['wh'])
        off_loss = reg_l1(out['reg'], batch['rot_mask'], batch['ind'], batch['reg'])
        w = self.w
        loss = w[0] * hm_loss + w[1] * wh_loss + w[2] * off_loss + w[3] * depth_loss + w[4] * dim_loss + w[5] * orien_loss + w[6] * kept_loss
        stats = {'loss': loss, 'hm_loss': hm_loss, 'wh_loss': wh_loss, 'off_loss': off_loss, 'dim_loss': dim_loss,
                 'orien_loss': orien_loss, 'kept_loss': kept_loss}
        if self.cost_volume:
            stats['depth_loss'] = depth_loss
        return loss, stats


class ModelWithLoss(nn.Module):
    """stereoTrainer.ModelWithLoss (:30-64): GT RoIs -> network -> loss, with nothing staged through the host."""

    def __init__(self, model, loss, output_w=320, wh_scale=1.0, cost_volume=True):
        super().__init__()
        self.model, self.loss = model, loss
        self.output_w, self.wh_scale, self.cost_volume = output_w, wh_scale, cost_volume

    def forward(self, batch):
        target = gt_rois(batch, self.output_w, self.wh_scale)
        outputs = self.model(batch, self.cost_volume, target)
        loss, stats = self.loss(outputs, batch)
        return outputs[-1], loss, stats
