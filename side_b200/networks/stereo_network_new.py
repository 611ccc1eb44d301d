"""``stereo_network_new`` -- the voxel / PointNet depth head variant.

Mirror of ``src/lib/models/networks/stereo_network_new.py`` + ``submodule.py:101-169``: same class names, constructor
arguments, ``forward(batch, useCostVolume=True, target=None)`` signature and ``state_dict()`` keys
(``feature_extraction.*``, ``feaRuduce.*``, ``pointNet.*``, the heads).  What runs differently:

* ``get_voxel`` (:160-283): a Python loop over images and RoIs on the host with ~40 small torch ops per RoI and five
  device -> host copies per image.  Here one kernel (``ops.voxel_coords``) for the drop-in function, and the fused
  ``ops.voxel_volume`` inside ``forward``: projection of the 10 x 10 x 10 metric grid into both views, bilinear sampling of the
  64-channel reduced features, validity masking and the ``cat(L - R, L, R)`` in ONE launch (the reference: two
  ``F.grid_sample`` per image, two mask multiplies, a subtraction and a concatenation over [N, 64, 1000] tensors).
* ``get_proposal_shift`` (:46-158) is vectorised on the device (no per-image loop, no ``.cpu()``).
* backbone, DCN neck and heads are the shared B200 modules of ``stereo_network.py`` (tensor-core paths in inference).
* ``PointNetDetector``'s 1x1 ``Conv1d`` / ``Linear`` stack are plain GEMMs and stay on cuBLAS.
"""
import importlib
import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..decode import bbox_decode  # noqa: F401
_old = importlib.import_module(__package__ + ".stereo_network")    # the package re-exports the class under the module's name

input_w = 1280      # module constants of the reference (:18-19)
input_h = 384


def project_rect_to_image(pts_3d_rect, P):
    """Reference :34-44."""
    n = pts_3d_rect.shape[0]
    ones = torch.ones((n, 1), device=pts_3d_rect.device)
    pts_2d = torch.mm(torch.cat([pts_3d_rect, ones], dim=1), torch.transpose(P, 0, 1))
    return pts_2d / pts_2d[:, 2:]


def get_proposal_shift(left_boxes, right_boxes, depth_rate, fbs, trans_invs):
    """Reference :46-158, vectorised: ``(proposals_left [D,N,5], proposals_right [D,N,5], depth_bin [N,D])`` with RoIs grouped
    by image in ascending image index (the reference concatenates per-image blocks)."""
    dev = left_boxes.device
    order = torch.argsort(left_boxes[:, 0], stable=True)
    lb, rb = left_boxes[order].float(), right_boxes[order].float()
    bi = lb[:, 0].long().clamp(0, fbs.shape[0] - 1)
    ti = trans_invs.float()[bi]                                    # [N, 2, 3]
    fb = fbs.float().reshape(-1)[bi]
    rate = torch.tensor([float(i) / (depth_rate - 1) for i in range(depth_rate)], dtype=torch.float32, device=dev)
    ox = lambda x, y: x * ti[:, 0, 0] + y * ti[:, 0, 1] + ti[:, 0, 2]
    center_left = (ox(lb[:, 1], lb[:, 2]) + ox(lb[:, 3], lb[:, 4])) / 2
    center_right = (ox(rb[:, 1], rb[:, 2]) + ox(rb[:, 3], rb[:, 4])) / 2
    center_disp = center_left - center_right
    dmin = torch.clamp((fb / center_disp) - 12.5, min=1.0, max=90.0).view(-1, 1)
    dmax = torch.clamp((fb / center_disp) + 12.5, min=1.0, max=90.0).view(-1, 1)
    depth_bin = dmax - (dmax - dmin) * rate
    disp = fb.view(-1, 1) / depth_bin / 4                          # [N, D]
    pro_left = lb.unsqueeze(0).expand(depth_rate, -1, -1).contiguous()
    pro_right = pro_left.clone()
    pro_right[:, :, 1] = torch.clamp(lb[:, 1].unsqueeze(0) - disp.t(), min=0)
    pro_right[:, :, 3] = torch.clamp(lb[:, 3].unsqueeze(0) - disp.t(), min=0)
    return pro_left, pro_right, depth_bin


def get_voxel(left_boxes, right_boxes, p2s, p3s, fbs, depth_bins, trans, trans_invs):
    """Reference :160-283 -> the same seven tensors, RoIs in image-major order, one kernel."""
    order = torch.argsort(left_boxes[:, 0], stable=True)
    return ops.voxel_coords(left_boxes[order].float().contiguous(), right_boxes[order].float().contiguous(), p2s.float().contiguous(),
                            p3s.float().contiguous(), fbs.float(), trans.float().contiguous(), trans_invs.float().contiguous(),
                            depth_bins.float().contiguous(), input_h, input_w)


class PointNetfeat_strAM(nn.Module):
    """submodule.py:101-131."""

    def __init__(self, input_c):
        super().__init__()
        self.conv1 = nn.Conv1d(input_c, 256, 1)
        self.conv2 = nn.Conv1d(256, 512, 1)
        self.conv3 = nn.Conv1d(512, 1024, 1)
        self.conv4 = nn.Conv1d(1024, 1024, 1)
        self.bn1 = nn.BatchNorm1d(256)
        self.bn2 = nn.BatchNorm1d(512)
        self.bn3 = nn.BatchNorm1d(1024)
        self.bn4 = nn.BatchNorm1d(1024)
        self.strAM_2D = nn.Conv2d(1024, 1024, 3, 1, 1)

    def forward(self, x, res):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.relu(self.bn2(self.conv2(x)))
        x = self.bn3(self.conv3(x))
        cube = x.view(x.size(0), x.size(1), res, res, res)
        isp = torch.sigmoid(self.strAM_2D(torch.mean(cube, dim=3))).unsqueeze(3)
        isp = (isp.expand_as(cube) * cube).view(x.size(0), x.size(1), res * res * res)
        x = F.relu(self.bn4(self.conv4(isp))) + x
        return torch.max(x, 2, keepdim=True)[0].view(-1, 1024)


class PointNetDetector(nn.Module):
    """submodule.py:133-169."""

    def __init__(self, input_c):
        super().__init__()
        self.feat_all = PointNetfeat_strAM(input_c)
        self.fc1 = nn.Linear(1024, 512)
        self.fc2 = nn.Linear(512, 256)
        self.depth = nn.Linear(256, 1)
        self.dropout = nn.Dropout(p=0.3)
        self.bn1 = nn.BatchNorm1d(512)
        self.bn2 = nn.BatchNorm1d(256)
        self.relu = nn.ReLU()
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2. / n))
            elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.bias.data.zero_()

    def forward(self, input_data, res):
        x = self.fc1(self.feat_all(input_data, res))
        x = F.relu(x if x.shape[0] <= 1 else self.bn1(x))
        x = self.dropout(self.fc2(x))
        x = F.relu(x if x.shape[0] <= 1 else self.bn2(x))
        return self.depth(x)


def fill_reduce_weights(layers):
    for m in layers.modules():
        if isinstance(m, nn.Conv2d):
            n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2. / n))
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


class stereo_network(_old.stereo_network):
    """Reference :300-463.  Shares backbone / neck / heads (and their tensor-core inference paths) with the cost-volume
    variant; the depth head is the voxel sampler + PointNetDetector."""

    def __init__(self, base_name, heads, pretrained, down_ratio, final_kernel, last_level, head_conv, out_channel=0):
        super().__init__(base_name, heads, pretrained, down_ratio, final_kernel, last_level, head_conv, out_channel)
        cf = self.feature_extraction.channels[self.first_level]
        del self.depth_estimator
        self.roiSize = 20
        self.feaRuduce = nn.Sequential(nn.Conv2d(cf, 64, kernel_size=3, padding=1, stride=1), nn.BatchNorm2d(64),
                                       nn.ReLU(inplace=True))
        self.pointNet = PointNetDetector(input_c=192)
        for head in self.heads:                        # registration order of the reference: ..., feaRuduce, pointNet, heads
            self._modules[head] = self._modules.pop(head)

    def forward(self, batch, useCostVolume=True, target=None):
        left, right = batch['input'], batch['input_right']
        imgfea_left, imgfea_right = self._features(left, right)
        if self._heads_tc_ok(imgfea_left):
            z = self._heads_tc(imgfea_left, imgfea_right)
        else:
            z, both = {}, None
            for head in self.heads:
                if head in self.left_only:
                    z[head] = self.__getattr__(head)(imgfea_left)
                else:
                    if both is None:
                        both = torch.cat((imgfea_left, imgfea_right), 1)
                    z[head] = self.__getattr__(head)(both)
        if useCostVolume:
            dev = left.device
            fb, p2, p3 = batch['fb'].to(dev), batch['p2'].to(dev), batch['p3'].to(dev)
            trans, trans_inv = batch['trans'].to(dev), batch['trans_inv'].to(dev)
            feaL = self.feaRuduce(imgfea_left)
            feaR = self.feaRuduce(imgfea_right)
            if target is not None:
                bbox_keep, bbox_right_keep, bboxShape = target
            else:
                bbox_keep, bbox_right_keep, bboxShape = bbox_decode(z['hm'], z['wh'], z['reg'])
            batch_size, max_obj = int(bboxShape[0]), int(bboxShape[1])
            depth = torch.zeros((batch_size, max_obj, 1), dtype=torch.float32, device=dev)
            if bbox_keep.shape[0] != 0:
                bl, br = bbox_keep.to(dev, torch.float32), bbox_right_keep.to(dev, torch.float32)
                order = torch.argsort(bl[:, 0], stable=True)        # RoIs grouped by image, as get_voxel emits them
                bl, br = bl[order].contiguous(), br[order].contiguous()
                voxel, depth_ori = ops.voxel_volume(feaL.contiguous(), feaR.contiguous(), bl, br, p2, p3, fb, trans, trans_inv,
                                                    input_h, input_w)
                disp = self.pointNet(voxel.reshape(voxel.shape[0], voxel.shape[1], -1), res=10)
                bi = bl[:, 0].long()
                onehot = bi.unsqueeze(1) == torch.arange(batch_size, device=dev).unsqueeze(0)
                slot = (torch.cumsum(onehot, 0) - 1).gather(1, bi.clamp(0, batch_size - 1).unsqueeze(1)).squeeze(1)
                depth = depth.index_put((bi, slot, torch.zeros_like(bi)), depth_ori + disp[:, 0])
            z.update({"depth": depth})
        return [z]


def get_pose_net(num_layers, heads, head_conv=256, down_ratio=4, pretrained=None):
    """Reference :466-473."""
    return stereo_network('dla{}'.format(num_layers), heads, pretrained=pretrained, down_ratio=down_ratio, final_kernel=1,
                          last_level=5, head_conv=head_conv)
