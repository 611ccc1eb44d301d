"""DLA-34 backbone + deformable up-sampling neck, state-dict compatible with the reference's
``feature_extraction_dla34.py`` (module tree ``base.*``, ``dla_up.ida_{0,1,2}.{proj,up,node}_k.*``,
``ida_up.{proj,up,node}_k.*``).

Only ``DeformConv.conv`` (the 16 DCN layers per view) is on the hot path and runs on
libside_b200.so; the DLA base, BatchNorms and the depth-wise bilinear ``ConvTranspose2d`` stay cuDNN/ATen
exactly as in the reference (out of scope).  No pretrained download happens here
(reference quirk Q4): ``pretrained`` may be a local ``.pth`` path or falsy.
"""
import math

import numpy as np
import torch
from torch import nn

from .. import ops
from ..conv_train import TCBatchNorm2d, TCConv2d
from ..dcn_v2 import DCN

BN_MOMENTUM = 0.1


def _bn(c):
    return TCBatchNorm2d(c, momentum=BN_MOMENTUM)


class BasicBlock(nn.Module):
    def __init__(self, inplanes, planes, stride=1, dilation=1):
        super().__init__()
        self.conv1 = TCConv2d(inplanes, planes, 3, stride=stride, padding=dilation, bias=False, dilation=dilation)
        self.bn1 = _bn(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = TCConv2d(planes, planes, 3, stride=1, padding=dilation, bias=False, dilation=dilation)
        self.bn2 = _bn(planes)
        self.stride = stride

    def forward(self, x, residual=None):
        residual = x if residual is None else residual
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        y += residual
        return self.relu(y)


class Root(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, residual):
        super().__init__()
        self.conv = TCConv2d(in_channels, out_channels, 1, stride=1, bias=False, padding=(kernel_size - 1) // 2)
        self.bn = _bn(out_channels)
        self.relu = nn.ReLU(inplace=True)
        self.residual = residual

    def forward(self, *xs):
        y = self.bn(self.conv(torch.cat(xs, 1)))
        if self.residual:
            y += xs[0]
        return self.relu(y)


class Tree(nn.Module):
    def __init__(self, levels, block, in_channels, out_channels, stride=1, level_root=False, root_dim=0,
                 root_kernel_size=1, dilation=1, root_residual=False):
        super().__init__()
        if root_dim == 0:
            root_dim = 2 * out_channels
        if level_root:
            root_dim += in_channels
        if levels == 1:
            self.tree1 = block(in_channels, out_channels, stride, dilation=dilation)
            self.tree2 = block(out_channels, out_channels, 1, dilation=dilation)
            self.root = Root(root_dim, out_channels, root_kernel_size, root_residual)
        else:
            self.tree1 = Tree(levels - 1, block, in_channels, out_channels, stride, root_dim=0,
                              root_kernel_size=root_kernel_size, dilation=dilation, root_residual=root_residual)
            self.tree2 = Tree(levels - 1, block, out_channels, out_channels, root_dim=root_dim + out_channels,
                              root_kernel_size=root_kernel_size, dilation=dilation, root_residual=root_residual)
        self.level_root = level_root
        self.root_dim = root_dim
        self.levels = levels
        self.downsample = nn.MaxPool2d(stride, stride=stride) if stride > 1 else None
        self.project = None
        if in_channels != out_channels:
            self.project = nn.Sequential(TCConv2d(in_channels, out_channels, 1, stride=1, bias=False), _bn(out_channels))

    def forward(self, x, residual=None, children=None):
        children = [] if children is None else children
        bottom = self.downsample(x) if self.downsample else x
        residual = self.project(bottom) if self.project else bottom
        if self.level_root:
            children.append(bottom)
        x1 = self.tree1(x, residual)
        if self.levels == 1:
            return self.root(self.tree2(x1), x1, *children)
        children.append(x1)
        return self.tree2(x1, children=children)


class DLA(ops.PreparedStateOwner, nn.Module):
    def __init__(self, levels, channels, block=BasicBlock, residual_root=False):
        super().__init__()
        self.channels = channels
        self.base_layer = nn.Sequential(nn.Conv2d(3, channels[0], 7, stride=1, padding=3, bias=False), _bn(channels[0]),
                                        nn.ReLU(inplace=True))
        self.level0 = self._conv_level(channels[0], channels[0], levels[0])
        self.level1 = self._conv_level(channels[0], channels[1], levels[1], stride=2)
        self.level2 = Tree(levels[2], block, channels[1], channels[2], 2, level_root=False, root_residual=residual_root)
        self.level3 = Tree(levels[3], block, channels[2], channels[3], 2, level_root=True, root_residual=residual_root)
        self.level4 = Tree(levels[4], block, channels[3], channels[4], 2, level_root=True, root_residual=residual_root)
        self.level5 = Tree(levels[5], block, channels[4], channels[5], 2, level_root=True, root_residual=residual_root)

    @staticmethod
    def _conv_level(inplanes, planes, convs, stride=1, dilation=1):
        mods = []
        for i in range(convs):
            mods += [TCConv2d(inplanes, planes, 3, stride=stride if i == 0 else 1, padding=dilation, bias=False,
                               dilation=dilation), _bn(planes), nn.ReLU(inplace=True)]
            inplanes = planes
        return nn.Sequential(*mods)

    def forward(self, x):
        ys = []
        stem = self.direct_stem and x.is_cuda and not self.training and not torch.is_grad_enabled()
        if stem and self.stem_block_convs and self.skip_stem_outputs and self._stem_tc_ok(x):
            # levels 0 / 1 are not consumed downstream (first_level >= 2): the stem hands level 2 its channels-last operand pairs
            return [None, None] + self._levels_tc_from(self._stem_tc(x), x.shape[0])
        x = self._stem(self.base_layer)(x) if stem else self.base_layer(x)
        for i in range(6):
            if i == 2 and self._tc_ok(x):
                return ys + self._levels_tc(x)
            level = getattr(self, "level%d" % i)
            x = self._stem(level)(x) if (stem and i < 2) else level(x)
            ys.append(x)
        return ys

    # -- stem on the tensor cores: base_layer (SIMT, 3 input channels) writes fp16 pairs in the 2x2 space-to-depth channels-last
    #    layout [B, H/2, W/2, 64]; in that layout level0 (16 -> 16, 3x3) is a 3x3 BLOCK convolution 64 -> 64 and level1 (16 -> 32,
    #    3x3, stride 2) a 3x3 block convolution 64 -> 32 at level-1 resolution (its stride disappears), both with 64-channel
    #    k-blocks instead of the 16 channels that starve an MMA.  Weights are rearranged once (structural zeros included). ----
    stem_block_convs = True
    skip_stem_outputs = False      # set by feature_extraction_dla34 when the up path starts at level >= 2

    def _stem_tc_ok(self, x):
        B, C, H, W = x.shape
        l0, l1 = list(self.level0), list(self.level1)
        return (self.tensor_core and ops.get_tc_format() == "f16" and C == 3 and H % 32 == 0 and W % 256 == 0 and len(l0) == 3
                and len(l1) == 3 and tuple(self.base_layer[0].weight.shape) == (16, 3, 7, 7) and tuple(l0[0].weight.shape) == (16, 16, 3, 3)
                and tuple(l1[0].weight.shape) == (32, 16, 3, 3) and l1[0].stride == (2, 2) and l0[0].stride == (1, 1)
                and self._tc_ok_shape(B, 32, H // 2, W // 2))

    @staticmethod
    def _block_weights(w, stride):
        """w [Cout, 16, 3, 3] of a pixel-domain 3x3 convolution -> the 3x3 convolution over 2x2 pixel blocks it equals:
        stride 1: [4 * Cout, 64, 3, 3] (output channel (dyo * 2 + dxo) * Cout + o), stride 2: [Cout, 64, 3, 3]; input channel
        (dyi * 2 + dxi) * 16 + c; block tap (by + 1, bx + 1)."""
        Cout, Cin = w.shape[:2]
        outs = [(0, 0)] if stride == 2 else [(0, 0), (0, 1), (1, 0), (1, 1)]
        wb = w.new_zeros((len(outs) * Cout, 4 * Cin, 3, 3))
        for io, (dyo, dxo) in enumerate(outs):
            for ky in range(3):
                ry = dyo + ky - 1                       # input row relative to the block's first row
                by, dyi = ry // 2, ry % 2
                for kx in range(3):
                    rx = dxo + kx - 1
                    bx, dxi = rx // 2, rx % 2
                    ci = (dyi * 2 + dxi) * Cin
                    wb[io * Cout:(io + 1) * Cout, ci:ci + Cin, by + 1, bx + 1] = w[:, :, ky, kx]
        return wb

    def _stem_tc_state(self):
        mods = [self.base_layer[0], self.base_layer[1], self.level0[0], self.level0[1], self.level1[0], self.level1[1]]
        key = (ops.prep_epoch(),) + tuple((t.data_ptr(), t._version) for m in mods for t in list(m.parameters()) + list(m.buffers()))
        st = self.__dict__.get("_stem_tc_cache")
        if st is None or st[0] != key:
            def affine(bn):
                sc = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float()
                return sc.contiguous(), (bn.bias - bn.running_mean * sc).detach().float().contiguous()
            s0, h0 = affine(self.base_layer[1])
            s1, h1 = affine(self.level0[1])
            s2, h2 = affine(self.level1[1])
            w1 = self._block_weights(self.level0[0].weight.detach().float(), 1).unsqueeze(2)       # [64, 64, 1, 3, 3]
            w2 = self._block_weights(self.level1[0].weight.detach().float(), 2).unsqueeze(2)       # [32, 64, 1, 3, 3]
            wp1, wp2 = ops.conv_tc_prepare(w1, fmt="f16"), ops.conv_tc_prepare(w2, fmt="f16")
            wp1.cin_alg = wp2.cin_alg = 16         # algorithmic FLOPs (bench.py) are those of the 16-channel pixel-domain layers
            st = (key, (self.base_layer[0].weight.detach().float().contiguous(), s0, h0),
                  (wp1, 64, s1.repeat(4).contiguous(), h1.repeat(4).contiguous()), (wp2, 32, s2, h2))
            self.__dict__["_stem_tc_cache"] = st
        return st[1:]

    def _stem_tc(self, x):
        """-> (full, hi, lo) of the level-1 output, channels-last [1, B, H/2, W/2, 32]: the triple level 2 starts from."""
        (w0, s0, h0), (wp1, c1, s1, h1), (wp2, c2, s2, h2) = self._stem_tc_state()
        hi, lo = ops.stem_conv_s2d(x, w0, s0, h0, relu=True)                                      # [B, 1, H/2, W/2, 64]
        _, hi, lo = ops.conv3d_tc(hi, lo, wp1, c1, ksize=(1, 3, 3), scale=s1, shift=h1, relu=True, full=False, split=True)
        full, hi, lo = ops.conv3d_tc(hi, lo, wp2, c2, ksize=(1, 3, 3), scale=s2, shift=h2, relu=True, full=True, split=True)
        B, _, H2, W2, C = hi.shape
        return full.view(1, B, H2, W2, C), hi.view(1, B, H2, W2, C), lo.view(1, B, H2, W2, C)

    # -- stem (base_layer, level0, level1) as direct fp32 convolutions with BatchNorm + ReLU folded in, inference only -----
    direct_stem = True

    def _stem(self, seq):
        def run(x):
            mods = list(seq)
            for j in range(0, len(mods), 3):                      # (Conv2d, BatchNorm2d, ReLU) triples
                conv, bn = mods[j], mods[j + 1]
                shape = (conv.in_channels, conv.out_channels, conv.kernel_size[0], conv.stride[0])
                if shape not in ((3, 16, 7, 1), (16, 16, 3, 1), (16, 32, 3, 2)) or conv.dilation[0] != 1:
                    x = mods[j + 2](bn(conv(x)))                  # not a DLA-34 stem shape: cuDNN
                    continue
                st = self.__dict__.setdefault("_stem_cache", {})
                key = (ops.prep_epoch(), bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
                       bn.weight.data_ptr())
                ent = st.get(id(bn))
                if ent is None or ent[0] != key:
                    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
                    ent = (key, scale, (bn.bias - bn.running_mean * scale).detach().float().contiguous())
                    st[id(bn)] = ent
                x = ops.stem_conv(x, conv.weight.detach(), ent[1], ent[2], stride=conv.stride[0], relu=True)
            return x
        return run

    # -- levels 2-5 (90 % of the base's FLOPs) on the tensor cores, inference only ---------------------------------
    # Every convolution of the four Trees is a tcgen05 implicit GEMM (ops.conv3d_tc, 3xFP16 / 3xTF32 operand pairs, fp32-class accuracy) on
    # channels-last activations with the batch as the box's depth axis: 3x3 and 1x1 kernels, stride-2 through TMA
    # element strides, eval-mode BatchNorm folded into the epilogue together with ReLU and the BasicBlock residual.
    tensor_core = True

    @staticmethod
    def _tiles(B, H, W):
        bw = 128
        while bw > 1 and W % bw:
            bw >>= 1
        bh = 128 // bw
        while bh > 1 and H % bh:
            bh >>= 1
        return bw * bh <= 128          # any batch: a depth box hanging over the batch is zero-filled by TMA (side_conv3d_tc_fwd)

    def _tc_ok(self, x):
        if not (self.tensor_core and x.is_cuda and not self.training and not torch.is_grad_enabled()):
            return False
        B, C, H, W = x.shape
        return self._tc_ok_shape(B, C, H, W)

    def _tc_ok_shape(self, B, C, H, W):
        """Level-1 output shape [B, C, H, W] from which levels 2-5 can run on the tensor cores."""
        if C % 32 or H % 16 or W % 16:
            return False
        return all(self._tiles(B, H >> k, W >> k) for k in (1, 2, 3, 4))

    def _tc_fold(self, conv, bn):
        st = self.__dict__.setdefault("_tc_cache", {})
        key = (ops.prep_epoch(), ops.get_tc_format(), conv.weight.data_ptr(), conv.weight._version, bn.weight._version,
               bn.bias._version, bn.running_mean._version, bn.running_var._version)
        ent = st.get(id(conv))
        if ent is None or ent[0] != key:
            scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
            shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
            ent = (key, ops.conv_tc_prepare(conv.weight.detach()), scale, shift)
            st[id(conv)] = ent
        return ent[1], ent[2], ent[3]

    def _conv_tc(self, t, conv, bn, relu, residual=None):
        """t = (full, hi, lo) channels-last [1, B, H, W, C] -> same triple after conv + folded BN (+ residual) (+ ReLU)."""
        wp, scale, shift = self._tc_fold(conv, bn)
        k = conv.kernel_size[0]
        return ops.conv3d_tc(t[1], t[2], wp, conv.out_channels, ksize=(1, k, k), scale=scale, shift=shift, relu=relu,
                             residual=residual, full=True, split=True, stride=conv.stride[0])

    def _block_tc(self, t, residual, blk):
        y = self._conv_tc(t, blk.conv1, blk.bn1, True)
        return self._conv_tc(y, blk.conv2, blk.bn2, "after", residual=residual)

    def _tree_tc(self, t, tree, children=None):
        children = [] if children is None else children
        bottom = ops.maxpool_hw2_cl(t[0], full=True, split=True) if tree.downsample is not None else t
        residual = self._conv_tc(bottom, tree.project[0], tree.project[1], False)[0] if tree.project is not None else bottom[0]
        if tree.level_root:
            children.append(bottom)
        if tree.levels == 1:
            x1 = self._block_tc(t, residual, tree.tree1)
            x2 = self._block_tc(x1, x1[0], tree.tree2)
            xs = [x2, x1] + children
            cat = (None, ops.cl_concat([c[1] for c in xs]), ops.cl_concat([c[2] for c in xs]))
            y = self._conv_tc(cat, tree.root.conv, tree.root.bn, "after" if tree.root.residual else True,
                              residual=xs[0][0] if tree.root.residual else None)
            return y
        x1 = self._tree_tc(t, tree.tree1)
        children.append(x1)
        return self._tree_tc(x1, tree.tree2, children=children)

    def _levels_tc(self, x):
        B, C, H, W = x.shape
        if ops.get_tc_format() == "f16":
            # fp16 pairs; the 32 input channels of level 2 are zero-padded to the 64-channel k-block (weights likewise)
            full, hi, lo = ops.ncdhw_to_cl_split(x.unsqueeze(2), want_full=True)
            Cp = hi.shape[-1]
            t = (full.view(1, B, H, W, Cp), hi.view(1, B, H, W, Cp), lo.view(1, B, H, W, Cp))
        else:
            hi, lo = ops.ncdhw_to_cl_split(x.unsqueeze(2))           # [B, 1, H, W, C] channels-last halves
            hi, lo = hi.view(1, B, H, W, C), lo.view(1, B, H, W, C)
            t = (hi + lo, hi, lo)                                    # hi + lo == x exactly
        return self._levels_tc_from(t, B)

    def _levels_tc_from(self, t, B):
        """Levels 2-5 from the channels-last (full, hi, lo) triple of the level-1 output [1, B, H, W, C]."""
        outs = []
        for i in range(2, 6):
            t = self._tree_tc(t, getattr(self, "level%d" % i))
            f = t[0]
            outs.append(ops.cl_to_nchw(f, B, f.shape[4], (f.shape[2], f.shape[3])))
        return outs

    def load_pretrained_model(self, path):
        """Loads ImageNet DLA weights from a LOCAL file (the reference downloads them, Q4)."""
        weights = torch.load(path, map_location="cpu")
        weights = {k: v for k, v in weights.items() if not k.startswith("fc.")}
        self.load_state_dict(weights, strict=False)


def dla34(pretrained=None, **kwargs):
    model = DLA([1, 1, 1, 2, 2, 1], [16, 32, 64, 128, 256, 512], block=BasicBlock, **kwargs)
    if isinstance(pretrained, str):
        model.load_pretrained_model(pretrained)
    return model


def fill_up_weights(up):
    """Bilinear kernel for the depth-wise ConvTranspose2d (reference feature_extraction_dla34.py:333-342)."""
    w = up.weight.data
    k = w.size(2)
    f = math.ceil(k / 2)
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    for i in range(k):
        for j in range(w.size(3)):
            w[0, 0, i, j] = (1 - math.fabs(i / f - c)) * (1 - math.fabs(j / f - c))
    w[1:, 0] = w[0, 0]


class DepthwiseUp(nn.ConvTranspose2d):
    """IDAUp.up_k: depth-wise bilinear-initialised ConvTranspose2d (reference :370-373), same parameters / state-dict
    key (``up_k.weight`` [o,1,2f,2f]); forward runs on side_dw_deconv_fwd instead of cuDNN's grouped direct kernel."""

    def forward(self, x):
        return ops.dw_deconv(x, self.weight, self.stride[0], self.padding[0])


class DeformConv(nn.Module):
    """DCN -> BN -> ReLU (reference :345-357).  In eval / no-grad mode BN and ReLU are folded into the
    DCN kernel's epilogue (SIDE_DCN_FUSE_AFFINE | SIDE_DCN_FUSE_RELU)."""

    def __init__(self, chi, cho):
        super().__init__()
        self.actf = nn.Sequential(_bn(cho), nn.ReLU(inplace=True))
        self.conv = DCN(chi, cho, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1)

    def forward(self, x):
        bn = self.actf[0]
        if not bn.training and not torch.is_grad_enabled():
            return self.conv(x, bn=bn, relu=True)
        return self.actf(self.conv(x))


class IDAUp(nn.Module):
    def __init__(self, o, channels, up_f):
        super().__init__()
        for i in range(1, len(channels)):
            c, f = channels[i], int(up_f[i])
            setattr(self, "proj_%d" % i, DeformConv(c, o))
            up = DepthwiseUp(o, o, f * 2, stride=f, padding=f // 2, output_padding=0, groups=o, bias=False)
            fill_up_weights(up)
            setattr(self, "up_%d" % i, up)
            setattr(self, "node_%d" % i, DeformConv(o, o))

    fuse_up_add = True     # inference: up_k + skip addition + layout change + operand split in one kernel (ops.idaup_fuse_cl)

    def forward(self, layers, startp, endp):
        for i in range(startp + 1, endp):
            k = i - startp
            proj, up, node = getattr(self, "proj_%d" % k), getattr(self, "up_%d" % k), getattr(self, "node_%d" % k)
            x, skip = proj(layers[i]), layers[i - 1]
            f, bn = up.stride[0], node.actf[0]
            if (self.fuse_up_add and not torch.is_grad_enabled() and not bn.training and ops.get_tc_format() == "f16"
                    and f in (2, 4, 8) and up.kernel_size[0] == 2 * f and up.padding[0] == f // 2 and node.conv._cl_ok(skip)
                    and tuple(skip.shape) == (x.shape[0], x.shape[1], x.shape[2] * f, x.shape[3] * f)):
                # everything between the two deformable convolutions in one pass; the up-sampled map and the sum never exist in NCHW
                full, hi, lo = ops.idaup_fuse_cl(x, up.weight, skip, f)
                layers[i] = node.conv.forward_prepared(full, hi, lo, tuple(skip.shape), bn=bn, relu=True)
                continue
            layers[i] = node(up(x) + skip)


class DLAUp(nn.Module):
    def __init__(self, startp, channels, scales, in_channels=None):
        super().__init__()
        self.startp = startp
        in_channels = list(channels) if in_channels is None else in_channels
        self.channels = channels
        channels = list(channels)
        scales = np.array(scales, dtype=int)
        for i in range(len(channels) - 1):
            j = -i - 2
            setattr(self, "ida_%d" % i, IDAUp(channels[j], in_channels[j:], scales[j:] // scales[j]))
            scales[j + 1:] = scales[j]
            in_channels[j + 1:] = [channels[j] for _ in channels[j + 1:]]

    def forward(self, layers):
        out = [layers[-1]]
        for i in range(len(layers) - self.startp - 1):
            getattr(self, "ida_%d" % i)(layers, len(layers) - i - 2, len(layers))
            out.insert(0, layers[-1])
        return out


class feature_extraction_dla34(nn.Module):
    """x [B,3,H,W] -> [B,64,H/4,W/4]  (reference :416-453)."""

    def __init__(self, base_name, pretrained, down_ratio, last_level, out_channel=0):
        super().__init__()
        assert down_ratio in [2, 4, 8, 16]
        assert base_name == "dla34", "only the canonical DLA-34 base is built"
        self.first_level = int(np.log2(down_ratio))
        self.last_level = last_level
        self.base = dla34(pretrained=pretrained)
        self.base.skip_stem_outputs = self.first_level >= 2       # DLAUp reads layers[first_level:] only
        self.channels = self.base.channels
        scales = [2 ** i for i in range(len(self.channels[self.first_level:]))]
        self.dla_up = DLAUp(self.first_level, self.channels[self.first_level:], scales)
        if out_channel == 0:
            out_channel = self.channels[self.first_level]
        self.ida_up = IDAUp(out_channel, self.channels[self.first_level:self.last_level],
                            [2 ** i for i in range(self.last_level - self.first_level)])

    def forward(self, x):
        x = self.dla_up(self.base(x))
        y = [x[i].clone() for i in range(self.last_level - self.first_level)]
        self.ida_up(y, 0, len(y))
        return y[-1]
