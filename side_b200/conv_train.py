"""Training-mode convolutions on the tensor cores (training side; BASELINE config #5).

The reference trains this network with cuDNN fp32 convolutions (modules/stereoTrainer.py:254-319 driving
stereo_network_old.py:139-171 and feature_extraction_dla34.py:31-95).  ``TCConv2d`` / ``TCConv3d`` are ``nn.Conv2d`` / ``nn.Conv3d``
(same parameters, same ``state_dict`` keys) whose forward and backward run on tcgen05 when autograd is recording:

* forward        ``side_conv3d_tc_fwd_f16`` on the (hi, lo) fp16 pairs of the input (the inference kernel, no BatchNorm folded --
                 training BatchNorm needs the batch statistics of the raw output);
* input gradient the same kernel on the spatially flipped, transposed weights (stride 1); grad_output is range-scaled by a power
                 of two taken from its absmax on the device and the scale is undone in the epilogue;
* weight gradient ``side_conv_wgrad_tc``: patch matrix copied from the saved input pairs + split-K tcgen05 GEMM.

Shapes the kernels cannot tile (stem layers, 12 x 40 maps in the weight gradient, stride-2 input gradients, 1-channel outputs)
keep the ATen / cuDNN implementation of that piece -- decided per call, silently only in that direction.
In ``eval()`` / ``no_grad`` the modules behave exactly like their parents (the inference fast paths read ``.weight`` directly).
"""
import torch
from torch import nn

from . import _lib, ops

enabled = True            # False: always the parent class' cuDNN path (A/B timing, debugging)
_unsupported = set()      # (piece, shape key) combinations the library refused once: not tried again


def _try(piece, key, fn):
    if (piece, key) in _unsupported:
        return None
    try:
        return fn()
    except RuntimeError as e:
        if "code -5" in str(e) or "code -1" in str(e) or "UNSUPPORTED" in str(e).upper():      # unsupported shape / bad argument
            _unsupported.add((piece, key))
            return None
        raise


def _pow2_scale(g, n_scale, n_inv):
    """Power-of-two s that brings max |g| into [2^10, 2^11) (1 where g is all zero; fp16 pairs resolve 2^-24 .. 65504), computed
    on the device: returns (s repeated n_scale times, 1 / s repeated n_inv times)."""
    buf = torch.empty((n_scale + n_inv + 1,), device=g.device, dtype=torch.float32)
    _lib.check(_lib.load().side_pow2_range_scale(g.data_ptr(), g.numel(), buf.data_ptr(), n_scale, buf[n_scale:].data_ptr(), n_inv,
                                                 buf[n_scale + n_inv:].data_ptr(), ops._stream()), "side_pow2_range_scale")
    return buf[:n_scale], buf[n_scale:n_scale + n_inv]


class _ConvTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, ksize):
        # x [N, Cin, D, H, W] (2-D convolutions arrive as [1, Cin, B, H, W]), weight [Cout, Cin, kd, kh, kw]
        hi, lo = ops.ncdhw_to_cl_split(x, fmt="f16")
        Cout = weight.shape[0]
        wp = ops.conv_tc_prepare(weight.detach(), fmt="f16", normalize=False)
        y, _, _ = ops.conv3d_tc(hi, lo, wp, Cout, ksize=ksize, full=True, split=False, stride=stride)
        ctx.save_for_backward(hi, lo, weight)
        ctx.meta = (tuple(x.shape), stride, ksize)
        N, Do, Ho, Wo, _ = y.shape
        return ops.cl_to_nchw(y, N, Cout, (Do, Ho, Wo))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        hi, lo, weight = ctx.saved_tensors
        xshape, stride, ksize = ctx.meta
        N, Cin, D, H, W = xshape
        Cout = weight.shape[0]
        gy = gy.contiguous()
        gx = gw = None
        key = (xshape, Cout, ksize, stride)
        if ctx.needs_input_grad[1]:
            gw = _try("wgrad", key, lambda: ops.conv_wgrad_tc(hi, lo, gy, Cin, ksize, stride))
            if gw is None:
                x = _pairs_to_nchw(hi, lo, Cin)
                gw = _aten_backward(x, weight, gy, stride, ksize, (False, True))[1]
        if ctx.needs_input_grad[0]:
            if stride == 1:
                gx = _try("dgrad", key, lambda: _dgrad_tc(gy, weight, ksize))
            if gx is None:
                x = _pairs_to_nchw(hi, lo, Cin)
                gx = _aten_backward(x, weight, gy, stride, ksize, (True, False))[0]
        return gx, gw, None, None


def _pairs_to_nchw(hi, lo, Cin):
    """The input back from its saved pairs (fallback paths only): hi + lo * 2^-11, channels-last -> NCDHW."""
    x = hi.float() + lo.float() * (1.0 / 2048.0)
    return x[..., :Cin].permute(0, 4, 1, 2, 3).contiguous()


def _aten_backward(x, weight, gy, stride, ksize, mask):
    pad = [(k - 1) // 2 for k in ksize]
    return torch.ops.aten.convolution_backward(gy, x, weight, None, [1, stride, stride], pad, [1, 1, 1], False, [0, 0, 0], 1,
                                               [mask[0], mask[1], False])


def _dgrad_tc(gy, weight, ksize):
    """grad_input of a stride-1 convolution = the convolution of grad_output with the flipped, transposed weights."""
    N, Cout, D, H, W = gy.shape
    Cin = weight.shape[1]
    s, inv = _pow2_scale(gy, N * D, Cin)
    hi, lo = ops.ncdhw_to_cl_split(gy, scale=s.view(N, D), fmt="f16")
    wt = weight.detach().flip(2, 3, 4).transpose(0, 1).contiguous()                       # [Cin, Cout, kd, kh, kw]
    wp = ops.conv_tc_prepare(wt, fmt="f16", normalize=False)
    gx, _, _ = ops.conv3d_tc(hi, lo, wp, Cin, ksize=ksize, scale=inv, full=True, split=False)          # shift NULL = 0
    return ops.cl_to_nchw(gx, N, Cin, (D, H, W))


def _eligible(x, conv, ksize, stride):
    if not (enabled and x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled() and ops.get_tc_format() == "f16"):
        return False
    Cout, Cin = conv.weight.shape[:2]
    return (conv.groups == 1 and all(d == 1 for d in conv.dilation) and conv.padding_mode == "zeros" and Cin % 32 == 0
            and Cout % 16 == 0 and (Cout <= 128 or Cout % 128 == 0) and ksize in ((1, 1, 1), (1, 3, 3), (3, 3, 3))
            and stride in (1, 2) and not (stride == 2 and ksize[0] != 1))


class TCConv2d(nn.Conv2d):
    def forward(self, x):
        ksize = (1,) + tuple(self.kernel_size)
        stride = self.stride[0]
        if (self.stride[0] == self.stride[1] and tuple(self.padding) == tuple((k - 1) // 2 for k in self.kernel_size)
                and _eligible(x, self, ksize, stride)):
            key = (tuple(x.shape), self.out_channels, ksize, stride)
            # [B, C, H, W] as [B, C, 1, H, W]: a view, and grad_output arrives as the [N, Cout, P] the weight gradient reads
            y = _try("fwd", key, lambda: _ConvTC.apply(x.unsqueeze(2), self.weight.unsqueeze(2), stride, ksize))
            if y is not None:
                y = y.squeeze(2)
                return y if self.bias is None else y + self.bias.view(1, -1, 1, 1)
        return super().forward(x)


class TCConv3d(nn.Conv3d):
    def forward(self, x):
        ksize = tuple(self.kernel_size)
        if (all(s == 1 for s in self.stride) and tuple(self.padding) == tuple((k - 1) // 2 for k in ksize)
                and _eligible(x, self, ksize, 1)):
            key = (tuple(x.shape), self.out_channels, ksize, 1)
            y = _try("fwd", key, lambda: _ConvTC.apply(x, self.weight, 1, ksize))
            if y is not None:
                return y if self.bias is None else y + self.bias.view(1, -1, 1, 1, 1)
        return super().forward(x)


# ----------------------------------------------------------------------------------------------------------------------------
# training-mode BatchNorm (side_bn_train_fwd / _bwd, csrc/batchnorm.cu)
# ----------------------------------------------------------------------------------------------------------------------------
class _BNTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, rm_ptr, rv_ptr, eps, momentum):
        # rm_ptr / rv_ptr: data pointers of the running statistics (buffers outside the graph, updated in place by the kernel)
        lib = _lib.load()
        x = x.contiguous()
        N, C = x.shape[:2]
        S = x.numel() // (N * C)
        y = torch.empty_like(x)
        stat = torch.empty((2, C), device=x.device, dtype=torch.float32)
        nws = lib.side_bn_train_ws_bytes(N, C, S)
        ws = torch.empty((nws // 8 + 1,), device=x.device, dtype=torch.float64)
        _lib.check(lib.side_bn_train_fwd(x.data_ptr(), ops._p(weight), ops._p(bias), rm_ptr, rv_ptr,
                                         y.data_ptr(), stat[0].data_ptr(), stat[1].data_ptr(), N, C, S, float(eps), float(momentum),
                                         ws.data_ptr(), nws, ops._stream()), "side_bn_train_fwd")
        ctx.save_for_backward(x, weight, stat)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        lib = _lib.load()
        x, weight, stat = ctx.saved_tensors
        gy = gy.contiguous()
        N, C = x.shape[:2]
        S = x.numel() // (N * C)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gwb = torch.empty((2, C), device=x.device, dtype=torch.float32)
        nws = lib.side_bn_train_ws_bytes(N, C, S)
        ws = torch.empty((nws // 8 + 1,), device=x.device, dtype=torch.float64)
        _lib.check(lib.side_bn_train_bwd(x.data_ptr(), gy.data_ptr(), ops._p(weight), stat[0].data_ptr(), stat[1].data_ptr(), ops._p(gx),
                                         gwb[0].data_ptr(), gwb[1].data_ptr(), N, C, S, ws.data_ptr(), nws, ops._stream()),
                   "side_bn_train_bwd")
        return gx, (gwb[0] if weight is not None else None), gwb[1], None, None, None, None


class _TCBatchNorm:
    """Mixin over nn.BatchNorm2d / nn.BatchNorm3d: train()-mode forward / backward through side_bn_train_* (batch statistics,
    running-statistics update and num_batches_tracked exactly as the parent); eval mode and unsupported shapes use the parent."""

    def forward(self, x):
        if (self.training and enabled and x.is_cuda and x.dtype == torch.float32 and self.momentum is not None
                and self.track_running_stats and self.affine and x.shape[0] * x.shape[1] <= 65535 and x.numel() > 0):
            self._check_input_dim(x)
            if self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
            return _BNTrain.apply(x, self.weight, self.bias, self.running_mean.data_ptr(), self.running_var.data_ptr(), self.eps,
                                  self.momentum)
        return super().forward(x)


class TCBatchNorm2d(_TCBatchNorm, nn.BatchNorm2d):
    pass


class TCBatchNorm3d(_TCBatchNorm, nn.BatchNorm3d):
    pass
