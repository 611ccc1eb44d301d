"""Drop-in replacements for the reference's ``DCNv2/dcn_v2.py`` module API, backed by libside_b200.so.

Same names, constructor arguments, parameter names / shapes / init and error behaviour as
``DCNv2`` (dcn_v2.py:57-94), ``DCN`` (:97-128) and ``dcn_v2_conv`` (:54), so existing
checkpoints (``...conv.weight``, ``...conv.bias``, ``...conv.conv_offset_mask.{weight,bias}``)
load unchanged.  ``ext_shim()`` additionally re-exposes ``dcn_v2_forward`` / ``dcn_v2_backward``
with the pybind signatures of ``DCNv2/src/vision.cpp:4-9`` so the reference's unmodified
``dcn_v2.py`` can run on top of this library.
"""
import math
import types

import torch
from torch import nn

from . import ops
from .ops import dcn_v2_conv  # noqa: F401  (re-export: same call signature as dcn_v2.py:54)


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class DCNv2(ops.PreparedStateOwner, nn.Module):
    """Modulated deformable conv taking explicit offset / mask (reference dcn_v2.py:57-94)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.deformable_groups = deformable_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        fan = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
        bound = 1.0 / math.sqrt(fan)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            self.bias.zero_()

    def forward(self, input, offset, mask):
        kk = self.kernel_size[0] * self.kernel_size[1]
        assert 2 * self.deformable_groups * kk == offset.shape[1]
        assert self.deformable_groups * kk == mask.shape[1]
        return dcn_v2_conv(input, offset, mask, self.weight, self.bias, self.stride, self.padding, self.dilation,
                           self.deformable_groups)


class DCN(DCNv2):
    """DCNv2 + its own zero-initialised offset/mask conv (reference dcn_v2.py:97-128).

    forward: ``om = conv_offset_mask(x)`` stays a cuDNN conv; the chunk / cat / sigmoid of the reference
    (4 elementwise launches) is folded into the deformable kernel, which reads ``om`` in place.
    With ``fuse_bn_relu(bn)`` set by ``DeformConv`` in eval mode the following BatchNorm + ReLU is folded
    into the epilogue as well.
    """

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, dilation=1, deformable_groups=1):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, deformable_groups)
        ch = self.deformable_groups * 3 * self.kernel_size[0] * self.kernel_size[1]
        self.conv_offset_mask = nn.Conv2d(self.in_channels, ch, kernel_size=self.kernel_size, stride=self.stride,
                                          padding=self.padding, bias=True)
        self.init_offset()

    def init_offset(self):
        with torch.no_grad():
            self.conv_offset_mask.weight.zero_()
            self.conv_offset_mask.bias.zero_()

    # -- inference fast path: offset conv on tcgen05 too, everything channels-last -----------------------------------
    tensor_core_offsets = True

    @staticmethod
    def _tiles(B, H, W):
        bw = 128
        while bw > 1 and W % bw:
            bw >>= 1
        bh = 128 // bw
        while bh > 1 and H % bh:
            bh >>= 1
        return bw * bh <= 128          # any batch: a depth box hanging over the batch is zero-filled by TMA (side_conv3d_tc_fwd)

    def _cl_ok(self, x):
        B, C, H, W = x.shape
        return (self.tensor_core_offsets and x.is_cuda and not torch.is_grad_enabled() and self.deformable_groups == 1
                and ops.get_dcn_precision() != "fp32" and self.kernel_size == (3, 3) and self.stride == (1, 1)
                and self.padding == (1, 1) and self.dilation == (1, 1) and C % 32 == 0 and self.out_channels % 16 == 0
                and self.out_channels <= 256 and self._tiles(B, H, W))

    def _om_weights(self):
        """conv_offset_mask as a 32-output-channel tcgen05 convolution (27 real channels, zero padding), bias in the epilogue."""
        c = self.conv_offset_mask
        key = (ops.prep_epoch(), ops.get_tc_format(), c.weight.data_ptr(), c.weight._version, c.bias._version)
        st = self.__dict__.get("_om_cache")
        if st is None or st[0] != key:
            w = torch.zeros((32,) + tuple(c.weight.shape[1:]), device=c.weight.device, dtype=torch.float32)
            w[:c.out_channels] = c.weight.detach()
            b = torch.zeros((32,), device=c.weight.device, dtype=torch.float32)
            b[:c.out_channels] = c.bias.detach()
            st = (key, ops.conv_tc_prepare(w), b)
            self.__dict__["_om_cache"] = st
        return st[1], st[2]

    def _forward_cl(self, input, bn, relu):
        B, C, H, W = input.shape
        full, hi, lo = ops.ncdhw_to_cl_split(input.unsqueeze(2), want_full=True)            # [B, 1, H, W, C]
        return self.forward_prepared(full, hi, lo, (B, C, H, W), bn, relu)

    def forward_prepared(self, full, hi, lo, shape, bn=None, relu=False):
        """The channels-last inference path from an input that is already channels-last and split: ``full`` fp32 and the operand
        pairs ``hi`` / ``lo`` [B, 1, H, W, C] (``ops.ncdhw_to_cl_split(x.unsqueeze(2), want_full=True)`` or ``ops.idaup_fuse_cl``);
        ``shape`` = (B, C, H, W) of the NCHW tensor they stand for.  Returns NCHW like ``forward``."""
        B, C, H, W = shape
        wp, bias = self._om_weights()
        om, _, _ = ops.conv3d_tc(hi.view(1, B, H, W, C), lo.view(1, B, H, W, C), wp, 32, ksize=(1, 3, 3), shift=bias,
                                 full=True, split=False)                                       # [1, B, H, W, 32]
        scale = shift = None
        if bn is not None:
            scale = (bn.weight * torch.rsqrt(bn.running_var + bn.eps)).contiguous()
            shift = (bn.bias - bn.running_mean * scale).contiguous()
        return ops.dcn_fwd_cl(full.view(B, H, W, C), om.view(B, H, W, 32), self.weight, self.bias, self.stride, self.padding,
                              self.dilation, scale=scale, shift=shift, relu=relu)

    def forward(self, input, bn=None, relu=False):
        if self._cl_ok(input) and (bn is None or not bn.training):
            return self._forward_cl(input, bn, relu)
        om = self.conv_offset_mask(input)
        if self.deformable_groups != 1:
            # generic path, literally the reference's sequence
            o1, o2, mask = torch.chunk(om, 3, dim=1)
            return dcn_v2_conv(input, torch.cat((o1, o2), dim=1), torch.sigmoid(mask), self.weight, self.bias,
                               self.stride, self.padding, self.dilation, self.deformable_groups)
        if bn is not None and not torch.is_grad_enabled():
            scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
            shift = bn.bias - bn.running_mean * scale
            return ops.dcn_fused_infer(input, om, self.weight, self.bias, self.stride, self.padding, self.dilation,
                                       scale=scale.contiguous(), shift=shift.contiguous(), relu=relu)
        if not torch.is_grad_enabled():
            return ops.dcn_fused_infer(input, om, self.weight, self.bias, self.stride, self.padding, self.dilation)
        return ops.dcn_fused(input, om, self.weight, self.bias, self.stride, self.padding, self.dilation)


def ext_shim():
    """A module object with the 2 on-path functions of the reference's pybind ``_ext``
    (DCNv2/src/vision.cpp:5-6; argument order of DCNv2/src/dcn_v2.h:9-23, 41-56)."""
    ext = types.ModuleType("_ext")

    def dcn_v2_forward(input, weight, bias, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                       dilation_h, dilation_w, deformable_group):
        if (kernel_h, kernel_w) != tuple(weight.shape[2:4]):
            raise RuntimeError("Input shape and kernel shape wont match: (%d x %d vs %d x %d)." %
                               (kernel_h, kernel_w, weight.shape[2], weight.shape[3]))
        return ops.dcn_forward_raw(input, ops._chk(offset, "offset"), ops._chk(mask, "mask"), weight,
                                   ops._chk(bias, "bias"), (stride_h, stride_w), (pad_h, pad_w),
                                   (dilation_h, dilation_w), deformable_group)

    def dcn_v2_backward(input, weight, bias, offset, mask, grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                        pad_w, dilation_h, dilation_w, deformable_group):
        if not input.is_contiguous():
            raise RuntimeError("input tensor has to be contiguous")
        if not weight.is_contiguous():
            raise RuntimeError("weight tensor has to be contiguous")
        gx, go, gm, gw, gb = ops.dcn_backward_raw(input, ops._chk(offset, "offset"), ops._chk(mask, "mask"), weight,
                                                  grad_output, (stride_h, stride_w), (pad_h, pad_w),
                                                  (dilation_h, dilation_w), deformable_group)
        return [gx, go, gm, gw, gb]

    ext.dcn_v2_forward = dcn_v2_forward
    ext.dcn_v2_backward = dcn_v2_backward
    return ext
