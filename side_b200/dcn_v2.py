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


class DCNv2(nn.Module):
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

    def forward(self, input, bn=None, relu=False):
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
