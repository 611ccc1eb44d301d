"""Torch-facing wrappers of the C ABI (``include/side_b200.h``): tensors in, tensors out.

PyTorch is only plumbing here -- it owns device memory and the current stream; every
operator below is one call into ``libside_b200.so``.  CPU tensors are rejected with
``RuntimeError`` (the reference's ``_ext`` does the same: DCNv2/src/dcn_v2.h:38).
"""
import torch

from . import _lib

_F32 = torch.float32


# Prepared state (swizzled weight tiles, folded BatchNorm) is cached per module under keys built from parameter addresses and
# version counters.  Two events change a module's numbers WITHOUT touching those: nn.Module._apply (.cpu() / .cuda() swap
# `param.data`, and the allocator may hand the old address back) and a training-mode BatchNorm forward (torch.batch_norm updates
# running_mean / running_var without bumping their version counters).  Modules that own a cache call invalidate_prepared()
# from train() / _apply(); every cache key carries the epoch.
_prep_epoch = 0


def invalidate_prepared():
    global _prep_epoch
    _prep_epoch += 1


def prep_epoch():
    return _prep_epoch


class PreparedStateOwner:
    """Mixin for nn.Modules that cache prepared weights: drops the caches on train() / eval() / .to() / .cuda() / .cpu()."""

    def train(self, mode=True):
        invalidate_prepared()
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        invalidate_prepared()
        return super()._apply(fn, *args, **kwargs)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, name, dtype=_F32):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("side_b200: %s must be a CUDA tensor -- not implemented on the CPU" % name)
    if t.dtype != dtype:
        raise RuntimeError("side_b200: %s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def _p(t):
    return None if t is None else t.data_ptr()


def _pair(v):
    return (int(v[0]), int(v[1])) if isinstance(v, (tuple, list)) else (int(v), int(v))


# ----------------------------------------------------------------------------------------------
# DCNv2
# ----------------------------------------------------------------------------------------------
PRECISIONS = {"fp32": _lib.DCN_PREC_FP32, "3xtf32": _lib.DCN_PREC_3XTF32, "tf32": _lib.DCN_PREC_TF32,
              "3xfp16": _lib.DCN_PREC_3XFP16}
_default_precision = "3xtf32"


def set_dcn_precision(name):
    """Selects the arithmetic of the DCN contraction: '3xtf32' (default: tcgen05 with the exact hi/lo split, fp32-class
    accuracy <= 1e-4 rel.; shapes the tensor-core kernel cannot tile -- Cin % 32, Cout % 16, Cout > 256, dg > 1 -- run the fp32
    SIMT kernel), '3xfp16' (tcgen05 kind::f16 on fp16 (hi, lo) pairs: the same 22 significand bits at twice the MMA rate and
    half the operand bytes, valid while |x|, |w| stay inside fp16's range -- watched by the range guard, see tc_range_status;
    Cin % 64 != 0 runs as 3xtf32), 'fp32' (always SIMT FMA) or 'tf32' (tcgen05 single pass, ~1e-3 relative, opt-in)."""
    global _default_precision
    if name not in PRECISIONS:
        raise ValueError("unknown DCN precision %r" % (name,))
    _default_precision = name


def get_dcn_precision():
    return _default_precision


def dcn_forward_raw(x, offset_t, mask_t, weight, bias, stride, padding, dilation, dg, *, offset_bs=0, mask_bs=0,
                    flags=0, scale=None, shift=None, offset_ptr=None, mask_ptr=None, precision=None):
    """One call of side_dcn_fwd.  ``offset_ptr`` / ``mask_ptr`` override the tensors' data_ptr (aliasing
    into the 27-channel conv_offset_mask output)."""
    lib = _lib.load()
    x = _chk(x, "input")
    weight = _chk(weight, "weight")
    B, Cin, H, W = x.shape
    Cout, Cin_w, kh, kw = weight.shape
    if Cin_w != Cin:
        raise RuntimeError("Input shape and kernel channels wont match: (%d vs %d)." % (Cin, Cin_w))
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    dh, dw = _pair(dilation)
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    flags |= PRECISIONS[precision or _default_precision]
    if (precision or _default_precision) == "3xfp16":
        _range_guard(x.device)
    y = torch.empty((B, Cout, Ho, Wo), device=x.device, dtype=_F32)
    nws = lib.side_dcn_fwd_ws_bytes(B, Cin, H, W, Cout, kh, kw, flags)
    ws = torch.empty((max(nws, 16),), device=x.device, dtype=torch.uint8)
    rc = lib.side_dcn_fwd(x.data_ptr(), offset_ptr or offset_t.data_ptr(), mask_ptr or mask_t.data_ptr(),
                          weight.data_ptr(), _p(bias), _p(scale), _p(shift), y.data_ptr(), B, Cin, H, W, Cout, kh, kw,
                          sh, sw, ph, pw, dh, dw, dg, offset_bs, mask_bs, flags, ws.data_ptr(), nws, _stream())
    _lib.check(rc, "side_dcn_fwd")
    return y


def dcn_backward_raw(x, offset_t, mask_t, weight, gy, stride, padding, dilation, dg, *, offset_bs=0, mask_bs=0,
                     flags=0, offset_ptr=None, mask_ptr=None, goffset_ptr=None, gmask_ptr=None, goffset=None,
                     gmask=None, ws_bytes=None):
    lib = _lib.load()
    x = _chk(x, "input")
    weight = _chk(weight, "weight")
    gy = _chk(gy, "grad_output")
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    dh, dw = _pair(dilation)
    gx = torch.empty_like(x)
    gw = torch.empty_like(weight)
    gb = torch.empty((Cout,), device=x.device, dtype=_F32)
    if goffset_ptr is None:
        goffset = torch.empty_like(offset_t)
        gmask = torch.empty_like(mask_t)
        goffset_ptr, gmask_ptr = goffset.data_ptr(), gmask.data_ptr()
    P = gy.shape[2] * gy.shape[3]
    per_sample = 4 * Cin * kh * kw * P
    want = lib.side_dcn_bwd_ws_bytes(B, Cin, H, W, Cout, kh, kw, flags)     # whole batch in one pass
    if ws_bytes is None:
        ws_bytes = want
        if want > (1 << 30):      # only large requests are checked against the free memory (the query costs a driver call)
            free, _ = torch.cuda.mem_get_info(x.device)
            ws_bytes = min(want, max(per_sample, int(free * 0.5)))
    ws = torch.empty((max(int(ws_bytes), per_sample),), device=x.device, dtype=torch.uint8)
    rc = lib.side_dcn_bwd(x.data_ptr(), offset_ptr or offset_t.data_ptr(), mask_ptr or mask_t.data_ptr(),
                          weight.data_ptr(), gy.data_ptr(), gx.data_ptr(), goffset_ptr, gmask_ptr, gw.data_ptr(),
                          gb.data_ptr(), B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dg, offset_bs, mask_bs, flags,
                          ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "side_dcn_bwd")
    return gx, goffset, gmask, gw, gb


class _DCNv2(torch.autograd.Function):
    """Drop-in for the reference autograd Function (DCNv2/dcn_v2.py:16-51)."""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias, stride, padding, dilation, deformable_groups):
        ctx.stride, ctx.padding, ctx.dilation = _pair(stride), _pair(padding), _pair(dilation)
        ctx.dg = int(deformable_groups)
        offset = _chk(offset, "offset")
        mask = _chk(mask, "mask")
        bias = _chk(bias, "bias") if bias is not None else None
        out = dcn_forward_raw(input, offset, mask, weight, bias, ctx.stride, ctx.padding, ctx.dilation, ctx.dg)
        ctx.save_for_backward(input, offset, mask, weight, bias)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        input, offset, mask, weight, bias = ctx.saved_tensors
        gx, go, gm, gw, gb = dcn_backward_raw(input, offset, mask, weight, grad_output, ctx.stride, ctx.padding,
                                              ctx.dilation, ctx.dg)
        return gx, go, gm, gw, (gb if bias is not None else None), None, None, None, None


dcn_v2_conv = _DCNv2.apply


class _DCNFused(torch.autograd.Function):
    """DCN.forward (dcn_v2.py:118-128) with the chunk/cat/sigmoid folded away: ``om`` is the raw 27-channel
    conv_offset_mask output; channels 0..17 are consumed as the interleaved offsets exactly as stored and
    channels 18..26 are mask logits (sigmoid fused in the kernel)."""

    @staticmethod
    def forward(ctx, input, om, weight, bias, stride, padding, dilation):
        ctx.stride, ctx.padding, ctx.dilation = _pair(stride), _pair(padding), _pair(dilation)
        om = _chk(om, "conv_offset_mask output")
        kk = weight.shape[2] * weight.shape[3]
        if om.shape[1] != 3 * kk:
            raise RuntimeError("conv_offset_mask must have %d channels" % (3 * kk))
        P = om.shape[2] * om.shape[3]
        ctx.kk, ctx.P = kk, P
        out = dcn_forward_raw(input, om, om, weight, bias, ctx.stride, ctx.padding, ctx.dilation, 1,
                              offset_bs=3 * kk * P, mask_bs=3 * kk * P, flags=_lib.DCN_MASK_IS_LOGIT,
                              mask_ptr=om.data_ptr() + 4 * 2 * kk * P)
        ctx.save_for_backward(input, om, weight, bias)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        input, om, weight, bias = ctx.saved_tensors
        kk, P = ctx.kk, ctx.P
        gom = torch.empty_like(om)
        gx, _, _, gw, gb = dcn_backward_raw(input, om, om, weight, grad_output, ctx.stride, ctx.padding, ctx.dilation, 1,
                                            offset_bs=3 * kk * P, mask_bs=3 * kk * P, flags=_lib.DCN_MASK_IS_LOGIT,
                                            mask_ptr=om.data_ptr() + 4 * 2 * kk * P, goffset_ptr=gom.data_ptr(),
                                            gmask_ptr=gom.data_ptr() + 4 * 2 * kk * P)
        return gx, gom, gw, (gb if bias is not None else None), None, None, None


dcn_fused = _DCNFused.apply


def dcn_fused_infer(input, om, weight, bias, stride, padding, dilation, scale=None, shift=None, relu=False):
    """Inference-only DCN (+ folded eval-mode BatchNorm + ReLU of DeformConv) in one kernel."""
    om = _chk(om, "conv_offset_mask output")
    kk = weight.shape[2] * weight.shape[3]
    P = om.shape[2] * om.shape[3]
    flags = _lib.DCN_MASK_IS_LOGIT
    if scale is not None:
        flags |= _lib.DCN_FUSE_AFFINE
    if relu:
        flags |= _lib.DCN_FUSE_RELU
    return dcn_forward_raw(input, om, om, weight, bias, stride, padding, dilation, 1, offset_bs=3 * kk * P,
                           mask_bs=3 * kk * P, flags=flags, scale=scale, shift=shift,
                           mask_ptr=om.data_ptr() + 4 * 2 * kk * P)


# ----------------------------------------------------------------------------------------------
# instance cost volume, gate, soft-argmin
# ----------------------------------------------------------------------------------------------
def proposal_shift(left, right, fb, D, x_clamp):
    lib = _lib.load()
    left, right, fb = _chk(left, "left_boxes"), _chk(right, "right_boxes"), _chk(fb, "fb")
    N = left.shape[0]
    pl = torch.empty((D, N, 5), device=left.device, dtype=_F32)
    pr = torch.empty((D, N, 5), device=left.device, dtype=_F32)
    db = torch.empty((N, D), device=left.device, dtype=_F32)
    _lib.check(lib.side_proposal_shift(left.data_ptr(), right.data_ptr(), fb.data_ptr(), N, fb.numel(), D,
                                       float(x_clamp), pl.data_ptr(), pr.data_ptr(), db.data_ptr(), _stream()),
               "side_proposal_shift")
    return pl, pr, db


USE_NHWC_GATHER = True   # False selects the NCHW gather kernel (same results, slower; kept for odd shapes / tests)


_tc_format = "f16"


def set_tc_format(fmt):
    """Operand format of the tensor-core convolutions: "tf32" (3xTF32) or "f16" (3xFP16: kind::f16 MMAs on fp16 (hi, lo * 2^11)
    pairs -- the same 22-bit operands and fp32 accumulation at twice the tensor rate; |activations| must stay below 65504)."""
    global _tc_format
    if fmt not in ("tf32", "f16"):
        raise ValueError("tc format must be 'tf32' or 'f16'")
    _tc_format = fmt


def get_tc_format():
    return _tc_format


# -- range guard of the fp16 operand pairs (include/side_b200.h: side_tc_range_guard) ---------------------------------------------
_range_blocks = {}      # device index -> int32[_RANGE_SLOTS] ring registered with the library
_RANGE_SLOTS = 1024
_RANGE_SAT_BITS, _RANGE_TINY_BITS = 0x477FE000, 0x36800000      # 65504.0f, 2^-18 as float bit patterns
TC_RANGE_SATURATED, TC_RANGE_UNDERFLOW = 1, 2


def _range_guard(device):
    """Registers (once per device) the ring of slots the kernels that write fp16 pairs leave max |x| of their tensor in."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    blk = _range_blocks.get(idx)
    if blk is None:
        blk = torch.zeros(_RANGE_SLOTS, device=torch.device("cuda", idx), dtype=torch.int32)
        with torch.cuda.device(idx):
            _lib.check(_lib.load().side_tc_range_guard(blk.data_ptr(), _RANGE_SLOTS), "side_tc_range_guard")
        _range_blocks[idx] = blk
    return blk


def tc_range_status(device=None, reset=True):
    """-> flags accumulated since the last reset: TC_RANGE_SATURATED (an activation tensor reached 65504: fp16 pairs invalid) |
    TC_RANGE_UNDERFLOW (a non-zero tensor stayed below 2^-18: pairs lose the 1e-4 bar).  One small device-to-host copy (syncs)."""
    idx = (torch.device(device).index if device is not None else None)
    idx = torch.cuda.current_device() if idx is None else idx
    blk = _range_blocks.get(idx)
    if blk is None:
        return 0
    hi = blk.max()                                                   # float bit patterns of non-negative values order as ints
    lo = torch.where(blk > 0, blk, torch.full_like(blk, 0x7FFFFFFF)).min()
    hi, lo = int(hi.item()), int(lo.item())
    flags = (TC_RANGE_SATURATED if hi >= _RANGE_SAT_BITS else 0) | (TC_RANGE_UNDERFLOW if lo < _RANGE_TINY_BITS else 0)
    if reset and hi:
        blk.zero_()
    return flags


VOL_BWD_FLAGS = 0   # tests / benchmarks: _lib.VOL_BWD_SCALAR forces the scalar-atomic backward kernel


class _InstCostVol(torch.autograd.Function):
    @staticmethod
    def forward(ctx, featL, featR, left, right, fb, valid, D, P, x_clamp, gate, fma=False, separable=False):
        lib = _lib.load()
        featL, featR = _chk(featL, "featL"), _chk(featR, "featR")
        left, right, fb = _chk(left, "left_boxes"), _chk(right, "right_boxes"), _chk(fb, "fb")
        if valid is not None:
            valid = _chk(valid, "valid", torch.uint8)
        B, C, H, W = featL.shape
        N = left.shape[0]
        cost = torch.empty((N, 3 * C, D, P, P), device=featL.device, dtype=_F32)
        depth_bin = torch.empty((N, D), device=featL.device, dtype=_F32)
        xc = torch.empty((N, D), device=featL.device, dtype=_F32) if gate else None
        flags = (_lib.VOL_GATE if gate else 0) | (_lib.VOL_FMA if fma else 0)
        if separable:
            flags |= _lib.VOL_SEPARABLE
            nws = lib.side_inst_costvol_fast_ws_bytes(B, C, H, W, N, D)
        else:
            nws = lib.side_inst_costvol_ws_bytes(B, C, H, W) if USE_NHWC_GATHER else 0
        ws = torch.empty((max(nws, 16),), device=featL.device, dtype=torch.uint8)
        _lib.check(lib.side_inst_costvol_fwd(featL.data_ptr(), featR.data_ptr(), left.data_ptr(), right.data_ptr(),
                                             fb.data_ptr(), _p(valid), cost.data_ptr(), depth_bin.data_ptr(), _p(xc),
                                             N, B, C, H, W, D, P, float(x_clamp), flags, ws.data_ptr() if nws else None,
                                             nws, _stream()),
                   "side_inst_costvol_fwd")
        ctx.save_for_backward(featL, featR, left, right, fb, valid)
        ctx.cfg = (D, P, float(x_clamp), flags & _lib.VOL_GATE)
        ctx.mark_non_differentiable(depth_bin)
        return cost, depth_bin

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gcost, _gdb):
        lib = _lib.load()
        featL, featR, left, right, fb, valid = ctx.saved_tensors
        D, P, x_clamp, flags = ctx.cfg
        B, C, H, W = featL.shape
        N = left.shape[0]
        gcost = _chk(gcost, "grad_cost")
        gL = torch.zeros_like(featL)
        gR = torch.zeros_like(featR)
        flags |= VOL_BWD_FLAGS
        nws = lib.side_inst_costvol_bwd_fast_ws_bytes(B, C, H, W, N, D, flags) if P == 16 and C % 8 == 0 else 0
        ws = torch.empty((nws,), device=featL.device, dtype=torch.uint8) if nws else None
        _lib.check(lib.side_inst_costvol_bwd_fast(featL.data_ptr(), featR.data_ptr(), left.data_ptr(), right.data_ptr(),
                                                  fb.data_ptr(), _p(valid), gcost.data_ptr(), gL.data_ptr(), gR.data_ptr(),
                                                  N, B, C, H, W, D, P, x_clamp, flags, _p(ws), nws, _stream()),
                   "side_inst_costvol_bwd_fast")
        return gL, gR, None, None, None, None, None, None, None, None, None, None


def inst_costvol(featL, featR, left, right, fb, D, P, x_clamp, gate=False, valid=None, fma=False, separable=False):
    """-> (cost [N,3C,D,P,P], depth_bin [N,D]); boxes [N,5] grouped by image in ascending b.
    Default: bit-identical to torchvision's RoIAlign loop.  fma=True: FMA-contracted bilinear taps (<= 1e-6 relative).
    separable=True: y interpolation shared by all D candidates of a RoI (<= 1e-5 relative, several times faster)."""
    return _InstCostVol.apply(featL, featR, left, right, fb, valid, int(D), int(P), float(x_clamp), bool(gate), bool(fma),
                              bool(separable))


def inst_costvol_ungated(featL, featR, left, right, fb, D, P, x_clamp, valid=None):
    """Inference-only single pass of the separable builder: -> (cost UNGATED [N,3C,D,P,P], depth_bin [N,D],
    xcross [N,D]); the consumer applies the gate (``ncdhw_to_cl_split(cost, scale=xcross)``)."""
    lib = _lib.load()
    featL, featR = _chk(featL, "featL"), _chk(featR, "featR")
    left, right, fb = _chk(left, "left_boxes"), _chk(right, "right_boxes"), _chk(fb, "fb")
    if valid is not None:
        valid = _chk(valid, "valid", torch.uint8)
    B, C, H, W = featL.shape
    N = left.shape[0]
    dev = featL.device
    cost = torch.empty((N, 3 * C, D, P, P), device=dev, dtype=_F32)
    depth_bin = torch.empty((N, D), device=dev, dtype=_F32)
    xc = torch.empty((N, D), device=dev, dtype=_F32)
    nws = lib.side_inst_costvol_fast_ws_bytes(B, C, H, W, N, D)
    ws = torch.empty((max(nws, 16),), device=dev, dtype=torch.uint8)
    _lib.check(lib.side_inst_costvol_fwd(featL.data_ptr(), featR.data_ptr(), left.data_ptr(), right.data_ptr(),
                                         fb.data_ptr(), _p(valid), cost.data_ptr(), depth_bin.data_ptr(), xc.data_ptr(),
                                         N, B, C, H, W, D, P, float(x_clamp), _lib.VOL_SEPARABLE | _lib.VOL_XCROSS,
                                         ws.data_ptr(), nws, _stream()), "side_inst_costvol_fwd")
    return cost, depth_bin, xc


def inst_costvol_cl_ok(C, D, P):
    """Shapes side_inst_costvol_fwd_cl is built for (the reference's: 32 reduced channels, 16x16 RoI bins)."""
    return C == 32 and P == 16 and 2 <= D <= 64


def inst_costvol_cl(featL, featR, left, right, fb, D, P, x_clamp, valid=None, gate=True, nhwc=False, diff=True):
    """Inference-only: the (gated) volume straight in its consumer's format -> (hi, lo fp16 [N, D, P, P, 3C], depth_bin [N, D],
    xcross [N, D]).  One pass over HBM; see include/side_b200.h side_inst_costvol_fwd_cl.  nhwc=True: the features are already
    channels-last [B, H, W, C].  diff=False (SIDE_VOL_NO_DIFF): only the L and R planes, [N, D, P, P, 2C] -- for a consumer that
    folded the L - R plane into its weights (cost_volume.aggregate_tc_pairs(..., folded=True))."""
    lib = _lib.load()
    featL, featR = _chk(featL, "featL"), _chk(featR, "featR")
    left, right, fb = _chk(left, "left_boxes"), _chk(right, "right_boxes"), _chk(fb, "fb")
    if valid is not None:
        valid = _chk(valid, "valid", torch.uint8)
    if nhwc:
        B, H, W, C = featL.shape
    else:
        B, C, H, W = featL.shape
    N = left.shape[0]
    dev = featL.device
    _range_guard(dev)
    hi = torch.empty((N, D, P, P, (3 if diff else 2) * C), device=dev, dtype=torch.float16)
    lo = torch.empty_like(hi)
    depth_bin = torch.empty((N, D), device=dev, dtype=_F32)
    xc = torch.empty((N, D), device=dev, dtype=_F32)
    nws = lib.side_inst_costvol_cl_ws_bytes(B, C, H, W)
    ws = torch.empty((nws,), device=dev, dtype=torch.uint8)
    _lib.check(lib.side_inst_costvol_fwd_cl(featL.data_ptr(), featR.data_ptr(), left.data_ptr(), right.data_ptr(), fb.data_ptr(),
                                            _p(valid), hi.data_ptr(), lo.data_ptr(), depth_bin.data_ptr(), xc.data_ptr(), N, B, C, H,
                                            W, D, P, float(x_clamp), (_lib.VOL_GATE if gate else 0) | (_lib.VOL_FEAT_NHWC if nhwc else 0) |
                                            (0 if diff else _lib.VOL_NO_DIFF),
                                            ws.data_ptr(), nws, _stream()),
               "side_inst_costvol_fwd_cl")
    return hi, lo, depth_bin, xc


class _XCrossGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cost, C):
        lib = _lib.load()
        cost = _chk(cost, "cost")
        N, C3, D, P, P2 = cost.shape
        if C3 != 3 * C or P != P2:
            raise RuntimeError("cost must be [N, 3*C, D, P, P]")
        out = torch.empty_like(cost)
        _lib.check(lib.side_xcross_gate_fwd(cost.data_ptr(), out.data_ptr(), None, N, C, D, P, _stream()),
                   "side_xcross_gate_fwd")
        ctx.save_for_backward(cost)
        ctx.C = C
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        lib = _lib.load()
        (cost,) = ctx.saved_tensors
        gout = _chk(gout, "grad_out")
        N, _, D, P, _ = cost.shape
        g = torch.empty_like(cost)
        _lib.check(lib.side_xcross_gate_bwd(cost.data_ptr(), gout.data_ptr(), g.data_ptr(), N, ctx.C, D, P, _stream()),
                   "side_xcross_gate_bwd")
        return g, None


def xcross_gate(cost, C):
    return _XCrossGate.apply(cost, int(C))


class _SoftArgmin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, depth_bin):
        lib = _lib.load()
        logits, depth_bin = _chk(logits, "logits"), _chk(depth_bin, "depth_bin")
        N, D, S, S2 = logits.shape
        if S != S2 or tuple(depth_bin.shape) != (N, D):
            raise RuntimeError("softargmin: logits [N,D,S,S], depth_bin [N,D] expected")
        depth = torch.empty((N,), device=logits.device, dtype=_F32)
        prob = torch.empty((N, D), device=logits.device, dtype=_F32)
        _lib.check(lib.side_softargmin_fwd(logits.data_ptr(), depth_bin.data_ptr(), depth.data_ptr(), prob.data_ptr(),
                                           N, D, S, _stream()), "side_softargmin_fwd")
        ctx.save_for_backward(prob, depth_bin, depth)
        ctx.S = S
        return depth

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        prob, depth_bin, depth = ctx.saved_tensors
        g = _chk(g, "grad_depth")
        N, D = prob.shape
        S = ctx.S
        gl = torch.empty((N, D, S, S), device=prob.device, dtype=_F32)
        gdb = torch.empty((N, D), device=prob.device, dtype=_F32)
        _lib.check(lib.side_softargmin_bwd(prob.data_ptr(), depth_bin.data_ptr(), depth.data_ptr(), g.data_ptr(),
                                           gl.data_ptr(), gdb.data_ptr(), N, D, S, _stream()), "side_softargmin_bwd")
        return gl, gdb


def softargmin(logits, depth_bin):
    """logits [N,D,S,S] (classify output squeezed), depth_bin [N,D] -> depth [N]."""
    return _SoftArgmin.apply(logits, depth_bin)


# ----------------------------------------------------------------------------------------------
# decode
# ----------------------------------------------------------------------------------------------
def bbox_decode_raw(heat, wh, reg, K=100, wh_scale=1.0, heat_is_logit=True):
    """-> dict(bbox [B,K,5], bbox_right [B,K,5], keep [B*K] uint8, slot [B,K] i32, count [B] i32,
    score [B,K], ind [B,K] i32, cls [B,K] i32).  Fixed shapes, no host sync."""
    lib = _lib.load()
    heat, wh, reg = _chk(heat, "heat"), _chk(wh, "wh"), _chk(reg, "reg")
    B, Cat, H, W = heat.shape
    dev = heat.device
    o = dict(bbox=torch.empty((B, K, 5), device=dev, dtype=_F32), bbox_right=torch.empty((B, K, 5), device=dev, dtype=_F32),
             keep=torch.empty((B * K,), device=dev, dtype=torch.uint8), slot=torch.empty((B, K), device=dev, dtype=torch.int32),
             count=torch.empty((B,), device=dev, dtype=torch.int32), score=torch.empty((B, K), device=dev, dtype=_F32),
             ind=torch.empty((B, K), device=dev, dtype=torch.int32), cls=torch.empty((B, K), device=dev, dtype=torch.int32))
    nws = lib.side_decode_ws_bytes(B, Cat, K)
    ws = torch.empty((nws,), device=dev, dtype=torch.uint8)
    flags = _lib.DECODE_HEAT_IS_LOGIT if heat_is_logit else 0
    _lib.check(lib.side_bbox_decode(heat.data_ptr(), wh.data_ptr(), reg.data_ptr(), o["bbox"].data_ptr(),
                                    o["bbox_right"].data_ptr(), o["keep"].data_ptr(), o["slot"].data_ptr(),
                                    o["count"].data_ptr(), o["score"].data_ptr(), o["ind"].data_ptr(), o["cls"].data_ptr(),
                                    B, Cat, H, W, K, float(wh_scale), flags, ws.data_ptr(), nws, _stream()),
               "side_bbox_decode")
    return o


def ddd_decode_raw(heat, kept, dim, orien, wh, reg, grid_size, K=40, heat_is_logit=False):
    lib = _lib.load()
    heat, kept, dim = _chk(heat, "heat"), _chk(kept, "kept"), _chk(dim, "dim")
    orien, wh, reg = _chk(orien, "orien"), _chk(wh, "wh"), _chk(reg, "reg")
    B, Cat, H, W = heat.shape
    if kept.shape[1] != 6 * grid_size:
        raise RuntimeError("kept must have 6*grid_size channels")
    dev = heat.device
    det = torch.empty((B, K, 6), device=dev, dtype=_F32)
    detr = torch.empty((B, K, 6), device=dev, dtype=_F32)
    info = torch.empty((B, K, 9), device=dev, dtype=_F32)
    nws = lib.side_decode_ws_bytes(B, Cat, K)
    ws = torch.empty((nws,), device=dev, dtype=torch.uint8)
    flags = _lib.DECODE_HEAT_IS_LOGIT if heat_is_logit else 0
    _lib.check(lib.side_ddd_decode(heat.data_ptr(), kept.data_ptr(), dim.data_ptr(), orien.data_ptr(), wh.data_ptr(),
                                   reg.data_ptr(), det.data_ptr(), detr.data_ptr(), info.data_ptr(), None, None, None, B,
                                   Cat, H, W, int(grid_size), K, flags, ws.data_ptr(), nws, _stream()), "side_ddd_decode")
    return det, detr, info


# ----------------------------------------------------------------------------------------------
# full-image volumes
# ----------------------------------------------------------------------------------------------
class _ConcatVolume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, L, R, D):
        lib = _lib.load()
        L, R = _chk(L, "left"), _chk(R, "right")
        B, C, H, W = L.shape
        vol = torch.empty((B, 2 * C, D, H, W), device=L.device, dtype=_F32)
        _lib.check(lib.side_concat_volume_fwd(L.data_ptr(), R.data_ptr(), vol.data_ptr(), B, C, H, W, D, _stream()),
                   "side_concat_volume_fwd")
        ctx.shape = (B, C, H, W, D)
        return vol

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        B, C, H, W, D = ctx.shape
        g = _chk(g, "grad_volume")
        gL = torch.empty((B, C, H, W), device=g.device, dtype=_F32)
        gR = torch.empty((B, C, H, W), device=g.device, dtype=_F32)
        _lib.check(lib.side_concat_volume_bwd(g.data_ptr(), gL.data_ptr(), gR.data_ptr(), B, C, H, W, D, _stream()),
                   "side_concat_volume_bwd")
        return gL, gR, None


def concat_volume(L, R, D):
    return _ConcatVolume.apply(L, R, int(D))


class _GwcVolume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, L, R, D, G):
        lib = _lib.load()
        L, R = _chk(L, "left"), _chk(R, "right")
        B, C, H, W = L.shape
        vol = torch.empty((B, G, D, H, W), device=L.device, dtype=_F32)
        _lib.check(lib.side_gwc_volume_fwd(L.data_ptr(), R.data_ptr(), vol.data_ptr(), B, C, H, W, D, G, _stream()),
                   "side_gwc_volume_fwd")
        ctx.save_for_backward(L, R)
        ctx.cfg = (D, G)
        return vol

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        L, R = ctx.saved_tensors
        D, G = ctx.cfg
        B, C, H, W = L.shape
        g = _chk(g, "grad_volume")
        gL = torch.empty_like(L)
        gR = torch.empty_like(R)
        _lib.check(lib.side_gwc_volume_bwd(L.data_ptr(), R.data_ptr(), g.data_ptr(), gL.data_ptr(), gR.data_ptr(), B, C, H,
                                           W, D, G, _stream()), "side_gwc_volume_bwd")
        return gL, gR, None, None


def gwc_volume(L, R, D, G):
    return _GwcVolume.apply(L, R, int(D), int(G))


# ----------------------------------------------------------------------------------------------
# depth-wise transposed convolution (IDAUp.up_k)
# ----------------------------------------------------------------------------------------------
class _DwDeconv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, stride, pad):
        lib = _lib.load()
        x, w = _chk(x, "input"), _chk(w, "weight")
        B, C, H, W = x.shape
        k = w.shape[-1]
        if tuple(w.shape) != (C, 1, k, k):
            raise RuntimeError("depth-wise deconv weight must be [C,1,k,k]")
        Ho, Wo = (H - 1) * stride - 2 * pad + k, (W - 1) * stride - 2 * pad + k
        y = torch.empty((B, C, Ho, Wo), device=x.device, dtype=_F32)
        _lib.check(lib.side_dw_deconv_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), B, C, H, W, k, stride, pad, _stream()),
                   "side_dw_deconv_fwd")
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, pad)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        lib = _lib.load()
        x, w = ctx.saved_tensors
        stride, pad = ctx.cfg
        gy = _chk(gy, "grad_output")
        B, C, H, W = x.shape
        gx, gw = torch.empty_like(x), torch.empty_like(w)
        _lib.check(lib.side_dw_deconv_bwd(x.data_ptr(), w.data_ptr(), gy.data_ptr(), gx.data_ptr(), gw.data_ptr(), B, C, H, W,
                                          w.shape[-1], stride, pad, _stream()), "side_dw_deconv_bwd")
        return gx, gw, None, None


def dw_deconv(x, w, stride, pad):
    """Depth-wise ConvTranspose2d: x [B,C,H,W], w [C,1,k,k]."""
    return _DwDeconv.apply(x, w, int(stride), int(pad))


# ----------------------------------------------------------------------------------------------
# aggregation network on the tensor cores (channels-last activations, tf32 hi/lo split)
# ----------------------------------------------------------------------------------------------
def conv_tc_prepare(weight, fmt=None, normalize=True):
    """nn.Conv3d / nn.Conv2d weight [Cout, Cin, *k] -> swizzled per-k-block tiles (hi, lo) for side_conv3d_tc_fwd.
    fmt="f16": fp16 pairs for the kind::f16 ("3xFP16") variant, returned as a float16 tensor."""
    lib = _lib.load()
    fmt = fmt or _tc_format
    weight = _chk(weight, "weight")
    Cout, Cin = weight.shape[:2]
    cin_alg = Cin                                    # unpadded input channels (algorithmic FLOP accounting in bench.py)
    taps = weight[0, 0].numel()
    nbytes = lib.side_conv_tc_weight_bytes(Cin, Cout, taps)
    if fmt == "f16":
        # fp16 pairs resolve 2^-24 .. 65504: weights far from O(1) are normalised by an exact power of two that conv3d_tc folds
        # back into the epilogue's per-channel scale (never taken by networks with sane initialisation: one host sync here)
        amax = float(weight.abs().max()) if normalize else 1.0       # normalize=False: no host sync (training steps)
        inv_scale = 1.0
        if amax > 0.0 and not (2.0 ** -8 <= amax <= 2.0 ** 8):
            import math
            e = math.floor(math.log2(amax))
            weight = weight * (2.0 ** -e)
            inv_scale = 2.0 ** e
        if Cin % 64:                                  # zero input channels up to the 64-wide k-block (matches ncdhw_to_cl_split)
            pad = 64 - Cin % 64
            weight = torch.cat((weight, weight.new_zeros((Cout, pad) + tuple(weight.shape[2:]))), 1).contiguous()
            Cin += pad
            nbytes = lib.side_conv_tc_weight_bytes(Cin, Cout, taps)
        wp = torch.empty((nbytes // 4,), device=weight.device, dtype=torch.float16)       # half the bytes of the tf32 tiles
        _lib.check(lib.side_conv_tc_prep_weights_f16(weight.data_ptr(), wp.data_ptr(), Cout, Cin, taps, _stream()),
                   "side_conv_tc_prep_weights_f16")
        wp.cin_alg = cin_alg
        wp.inv_scale = inv_scale
        return wp
    wp = torch.empty((nbytes // 4,), device=weight.device, dtype=_F32)
    _lib.check(lib.side_conv_tc_prep_weights(weight.data_ptr(), wp.data_ptr(), Cout, Cin, taps, _stream()),
               "side_conv_tc_prep_weights")
    return wp


def conv3d_tc(x_hi, x_lo, wp, Cout, ksize=(3, 3, 3), scale=None, shift=None, relu=False, residual=None, full=False,
              split=True, stride=1, pool=False):
    """x_hi, x_lo [N, D, H, W, Cin] -> (y, y_hi, y_lo) [N, D, H/stride, W/stride, Cout] (entries not requested are None).
    relu: False / True (before the residual) / "after" (after the residual, DLA BasicBlock).
    pool=True: MaxPool3d((1,2,2)) of the result fused into the epilogue (outputs [N, D, H/2, W/2, Cout])."""
    lib = _lib.load()
    f16 = x_hi.dtype == torch.float16          # operand format follows the activations: fp16 pairs -> kind::f16 kernel
    odt = torch.float16 if f16 else _F32
    x_hi, x_lo = _chk(x_hi, "x_hi", odt), _chk(x_lo, "x_lo", odt)
    if wp.dtype != odt:
        raise RuntimeError("conv3d_tc: weight tiles (%s) and activations (%s) use different operand formats" % (wp.dtype, odt))
    N, D, H, W, Cin = x_hi.shape
    dev = x_hi.device
    oshape = (N, D, H // stride // (2 if pool else 1), W // stride // (2 if pool else 1), Cout)
    y = torch.empty(oshape, device=dev, dtype=_F32) if full else None
    y_hi = torch.empty(oshape, device=dev, dtype=odt) if split else None
    y_lo = torch.empty(oshape, device=dev, dtype=odt) if split else None
    if residual is not None:
        residual = _chk(residual, "residual")
    inv = getattr(wp, "inv_scale", 1.0)
    if inv != 1.0:                                    # weights were normalised by a power of two (conv_tc_prepare)
        scale = torch.full((Cout,), inv, device=dev, dtype=_F32) if scale is None else scale * inv
    if f16 and split:
        _range_guard(dev)
    fn = lib.side_conv3d_tc_fwd_f16 if f16 else lib.side_conv3d_tc_fwd
    _lib.check(fn(x_hi.data_ptr(), x_lo.data_ptr(), wp.data_ptr(), _p(scale), _p(shift), _p(residual),
                                      _p(y), _p(y_hi), _p(y_lo), N, D, H, W, Cin, Cout, ksize[0], ksize[1], ksize[2],
                                      int(stride), (2 if relu == "after" else (1 if relu else 0)) | (4 if pool else 0), _stream()),
               "side_conv3d_tc_fwd")
    return y, y_hi, y_lo


def ncdhw_to_cl_split(x, scale=None, want_full=False, fmt=None):
    """x [N, C, D, H, W] (* scale [N, D] per depth slice) -> hi, lo [N, D, H, W, C] (and the unsplit copy first if want_full)."""
    lib = _lib.load()
    fmt = fmt or _tc_format
    x = _chk(x, "x")
    N, C = x.shape[:2]
    sp = tuple(x.shape[2:])
    S = 1
    for s in sp:
        S *= s
    D = 0
    if scale is not None:
        scale = _chk(scale, "scale")
        D = sp[0]
        if tuple(scale.shape) != (N, D):
            raise RuntimeError("scale must be [N, D]")
    if fmt == "f16":
        _range_guard(x.device)
        Cp = (C + 31) // 32 * 32                     # rows of 32-channel multiples (TMA zero-fills the tail of a 64-channel k-block)
        hi = torch.empty((N,) + sp + (Cp,), device=x.device, dtype=torch.float16)
        lo = torch.empty_like(hi)
        full = torch.empty((N,) + sp + (Cp,), device=x.device, dtype=_F32) if want_full else None
        _lib.check(lib.side_ncdhw_to_cl_split_f16(x.data_ptr(), _p(scale), _p(full), hi.data_ptr(), lo.data_ptr(), N, C, S, D, Cp,
                                                  _stream()), "side_ncdhw_to_cl_split_f16")
        return (full, hi, lo) if want_full else (hi, lo)
    hi = torch.empty((N,) + sp + (C,), device=x.device, dtype=_F32)
    lo = torch.empty_like(hi)
    full = torch.empty_like(hi) if want_full else None
    _lib.check(lib.side_ncdhw_to_cl_split(x.data_ptr(), _p(scale), _p(full), hi.data_ptr(), lo.data_ptr(), N, C, S, D, _stream()),
               "side_ncdhw_to_cl_split")
    return (full, hi, lo) if want_full else (hi, lo)


def idaup_fuse_cl(x, weight, skip, stride):
    """One IDAUp step between its two deformable convolutions (reference feature_extraction_dla34.py:380-386), inference, fp16
    pairs:  up_k(x) + skip  ->  (full fp32, hi, lo fp16) channels-last [B, 1, H*stride, W*stride, Cp] -- what
    ``ncdhw_to_cl_split((dw_deconv(x, weight, stride, stride // 2) + skip).unsqueeze(2), want_full=True)`` returns, in one pass."""
    lib = _lib.load()
    x, weight, skip = _chk(x, "x"), _chk(weight, "weight"), _chk(skip, "skip")
    B, C, H, W = x.shape
    Ho, Wo = H * stride, W * stride
    if tuple(skip.shape) != (B, C, Ho, Wo) or tuple(weight.shape) != (C, 1, 2 * stride, 2 * stride):
        raise RuntimeError("idaup_fuse_cl: skip must be [B, C, H*stride, W*stride] and weight [C, 1, 2*stride, 2*stride]")
    _range_guard(x.device)
    Cp = (C + 31) // 32 * 32
    full = torch.empty((B, 1, Ho, Wo, Cp), device=x.device, dtype=_F32)
    hi = torch.empty((B, 1, Ho, Wo, Cp), device=x.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    _lib.check(lib.side_idaup_fuse_cl_f16(x.data_ptr(), weight.data_ptr(), skip.data_ptr(), full.data_ptr(), hi.data_ptr(),
                                          lo.data_ptr(), B, C, H, W, int(stride), Cp, _stream()), "side_idaup_fuse_cl_f16")
    return full, hi, lo


def split_pairs(x, fmt=None):
    """Operand pairs (hi, lo) of a tensor that already has the consumer's (channels-last) layout, in the given operand format."""
    fmt = fmt or _tc_format
    if fmt != "f16":
        return tf32_split(x)
    lib = _lib.load()
    x = _chk(x, "x")
    _range_guard(x.device)
    hi = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    _lib.check(lib.side_f16_split(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream()), "side_f16_split")
    return hi, lo


def tf32_split(x):
    lib = _lib.load()
    x = _chk(x, "x")
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    _lib.check(lib.side_tf32_split(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream()), "side_tf32_split")
    return hi, lo


def gate_mul_split(y, gate, fmt=None):
    """y [N, D, H, W, C] * gate [N, D, W, C] (broadcast over H) -> hi, lo."""
    lib = _lib.load()
    fmt = fmt or _tc_format
    y, gate = _chk(y, "y"), _chk(gate, "gate")
    N, D, H, W, C = y.shape
    if tuple(gate.shape) != (N, D, W, C):
        raise RuntimeError("gate must be [N, D, W, C]")
    odt = torch.float16 if fmt == "f16" else _F32
    if fmt == "f16":
        _range_guard(y.device)
    hi, lo = torch.empty_like(y, dtype=odt), torch.empty_like(y, dtype=odt)
    fn = lib.side_gate_mul_split_f16 if fmt == "f16" else lib.side_gate_mul_split
    _lib.check(fn(y.data_ptr(), gate.data_ptr(), hi.data_ptr(), lo.data_ptr(), N, D, H, W, C, _stream()),
               "side_gate_mul_split")
    return hi, lo


def maxpool_hw2_cl(x, full=False, split=True, fmt=None):
    """MaxPool3d((1,2,2)) on channels-last x [N, D, H, W, C] -> (y, hi, lo) [N, D, H/2, W/2, C]."""
    lib = _lib.load()
    fmt = fmt or _tc_format
    x = _chk(x, "x")
    N, D, H, W, C = x.shape
    shp = (N, D, H // 2, W // 2, C)
    y = torch.empty(shp, device=x.device, dtype=_F32) if full else None
    odt = torch.float16 if fmt == "f16" else _F32
    if fmt == "f16" and split:
        _range_guard(x.device)
    hi = torch.empty(shp, device=x.device, dtype=odt) if split else None
    lo = torch.empty(shp, device=x.device, dtype=odt) if split else None
    fn = lib.side_maxpool_hw2_cl_f16 if fmt == "f16" else lib.side_maxpool_hw2_cl
    _lib.check(fn(x.data_ptr(), _p(y), _p(hi), _p(lo), N, D, H, W, C, _stream()), "side_maxpool_hw2_cl")
    return y, hi, lo


def conv3d_c1_cl(x, w):
    """x [N, D, H, W, C], w [1, C, 3, 3, 3] -> [N, D, H, W]."""
    lib = _lib.load()
    x, w = _chk(x, "x"), _chk(w, "w")
    N, D, H, W, C = x.shape
    if tuple(w.shape) != (1, C, 3, 3, 3):
        raise RuntimeError("conv3d_c1_cl: weight must be [1, C, 3, 3, 3]")
    out = torch.empty((N, D, H, W), device=x.device, dtype=_F32)
    _lib.check(lib.side_conv3d_c1_cl(x.data_ptr(), w.data_ptr(), out.data_ptr(), N, D, H, W, C, _stream()), "side_conv3d_c1_cl")
    return out


def stem_conv(x, weight, scale=None, shift=None, stride=1, relu=True):
    """Direct fp32 Conv2d(k, stride, pad (k-1)/2, no bias) + folded BatchNorm + ReLU for the DLA-34 stem shapes (NCHW)."""
    lib = _lib.load()
    x, weight = _chk(x, "x"), _chk(weight, "weight")
    B, Cin, H, W = x.shape
    Cout, _, k, _ = weight.shape
    p = (k - 1) // 2
    Ho, Wo = (H + 2 * p - k) // stride + 1, (W + 2 * p - k) // stride + 1
    y = torch.empty((B, Cout, Ho, Wo), device=x.device, dtype=_F32)
    _lib.check(lib.side_stem_conv_fwd(x.data_ptr(), weight.data_ptr(), _p(scale), _p(shift), y.data_ptr(), B, Cin, H, W, Cout, k,
                                      int(stride), 1 if relu else 0, _stream()), "side_stem_conv_fwd")
    return y


def stem_conv_s2d(x, weight, scale=None, shift=None, relu=True):
    """base_layer (3 -> 16, 7x7) + folded BatchNorm + ReLU -> fp16 operand pairs (hi, lo) [B, 1, H/2, W/2, 64] in the 2x2
    space-to-depth channels-last layout (channel = (dy * 2 + dx) * 16 + o), see side_stem_conv_fwd_s2d."""
    lib = _lib.load()
    x, weight = _chk(x, "x"), _chk(weight, "weight")
    B, Cin, H, W = x.shape
    if tuple(weight.shape) != (16, 3, 7, 7) or H % 2 or W % 2:
        raise RuntimeError("side_stem_conv_fwd_s2d failed (code -5): built for the 3 -> 16 7x7 base layer on even-sized images")
    _range_guard(x.device)
    hi = torch.empty((B, 1, H // 2, W // 2, 64), device=x.device, dtype=torch.float16)
    lo = torch.empty_like(hi)
    _lib.check(lib.side_stem_conv_fwd_s2d(x.data_ptr(), weight.data_ptr(), _p(scale), _p(shift), hi.data_ptr(), lo.data_ptr(), B, H, W,
                                          1 if relu else 0, _stream()), "side_stem_conv_fwd_s2d")
    return hi, lo


def cl_concat(tensors):
    """torch.cat(tensors, -1) for channels-last tensors of one dtype whose rows are multiples of 16 bytes (one 16-byte-copy kernel)."""
    lib = _lib.load()
    ts = [_chk(t, "tensor", tensors[0].dtype) for t in tensors]
    lead = ts[0].shape[:-1]
    es = ts[0].element_size()
    if len(ts) > 8 or any(t.shape[:-1] != lead or (t.shape[-1] * es) % 16 for t in ts):
        return torch.cat(ts, -1)
    rows = 1
    for v in lead:
        rows *= v
    out = torch.empty(tuple(lead) + (sum(t.shape[-1] for t in ts),), device=ts[0].device, dtype=ts[0].dtype)
    ptrs = (_lib._vp * len(ts))(*[t.data_ptr() for t in ts])
    nb = (_lib._i * len(ts))(*[t.shape[-1] * es for t in ts])
    _lib.check(lib.side_cl_concat(ptrs, nb, len(ts), out.data_ptr(), rows, _stream()), "side_cl_concat")
    return out


def cl_to_nchw(x, B, C, spatial, ld=None):
    """channels-last buffer [B, *spatial, ld] -> NCHW [B, C, *spatial] (ld > C: rows computed with padded channels, first C taken)."""
    lib = _lib.load()
    x = _chk(x, "x")
    HW = 1
    for v in spatial:
        HW *= v
    y = torch.empty((B, C) + tuple(spatial), device=x.device, dtype=_F32)
    _lib.check(lib.side_cl_to_nchw_ld(x.data_ptr(), int(ld or C), y.data_ptr(), B, C, HW, _stream()), "side_cl_to_nchw")
    return y


def dcn_fwd_cl(x_nhwc, om_cl, weight, bias, stride, padding, dilation, scale=None, shift=None, relu=False, precision=None):
    """Inference DCN on channels-last inputs: x_nhwc [B, H, W, Cin], om_cl [B, Ho, Wo, ld] (offsets then mask logits per pixel)
    -> y [B, Cout, Ho, Wo].  tcgen05 precisions only."""
    lib = _lib.load()
    x_nhwc, om_cl, weight = _chk(x_nhwc, "x_nhwc"), _chk(om_cl, "om_cl"), _chk(weight, "weight")
    B, H, W, Cin = x_nhwc.shape
    Cout, _, kh, kw = weight.shape
    sh, sw = _pair(stride); ph, pw = _pair(padding); dh, dw = _pair(dilation)
    Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    prec = precision or _default_precision
    flags = PRECISIONS["3xtf32" if prec == "fp32" else prec]
    if prec == "3xfp16":
        _range_guard(x_nhwc.device)
    if scale is not None:
        flags |= _lib.DCN_FUSE_AFFINE
    if relu:
        flags |= _lib.DCN_FUSE_RELU
    y = torch.empty((B, Cout, Ho, Wo), device=x_nhwc.device, dtype=_F32)
    nws = lib.side_dcn_fwd_ws_bytes(B, Cin, H, W, Cout, kh, kw, flags)
    ws = torch.empty((max(nws, 16),), device=x_nhwc.device, dtype=torch.uint8)
    _lib.check(lib.side_dcn_fwd_cl(x_nhwc.data_ptr(), om_cl.data_ptr(), om_cl.shape[-1], weight.data_ptr(), _p(bias), _p(scale),
                                   _p(shift), y.data_ptr(), B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, flags,
                                   ws.data_ptr(), nws, _stream()), "side_dcn_fwd_cl")
    return y


# ----------------------------------------------------------------------------------------------
# stereo_network_new: instance voxel volume
# ----------------------------------------------------------------------------------------------
voxel_align_corners = False      # grid_sample convention of the voxel sampler; False = torch >= 1.3 default


def voxel_coords(left, right, p2, p3, fb, trans, trans_inv, depth_bins, input_h=384, input_w=1280):
    """get_voxel (stereo_network_new.py:160-283) on the device: returns the reference's seven tensors
    (norm_coord_imgs, valids, norm_coord_left_imgs2ds, valids_left, norm_coord_right_imgs2ds, valids_right, depth_ori)."""
    lib = _lib.load()
    args = [_chk(t.to(_F32) if t.is_cuda else t, n) for t, n in ((left, "left_boxes"), (right, "right_boxes"), (p2, "p2"), (p3, "p3"),
                                                                 (fb.reshape(-1), "fb"), (trans, "trans"), (trans_inv, "trans_inv"),
                                                                 (depth_bins, "depth_bins"))]
    left, right, p2, p3, fb, trans, trans_inv, depth_bins = args
    N, B, D = left.shape[0], fb.shape[0], depth_bins.shape[1]
    dev = left.device
    e = lambda *s: torch.empty(s, device=dev, dtype=_F32)
    norm3, valid3, normL, validL = e(N, 10, 10, 10, 3), e(N, 10, 10, 10), e(N, 10, 10, 10, 2), e(N, 10, 10, 10)
    normR, validR, dori = e(N, 10, 10, 10, 2), e(N, 10, 10, 10), e(N)
    _lib.check(lib.side_voxel_coords(left.data_ptr(), right.data_ptr(), p2.data_ptr(), p3.data_ptr(), fb.data_ptr(),
                                     trans.data_ptr(), trans_inv.data_ptr(), depth_bins.data_ptr(), N, B, D, 0, 0, int(input_h),
                                     int(input_w), norm3.data_ptr(), valid3.data_ptr(), normL.data_ptr(), validL.data_ptr(),
                                     normR.data_ptr(), validR.data_ptr(), dori.data_ptr(), _stream()), "side_voxel_coords")
    return norm3, valid3, normL, validL, normR, validR, dori


class _VoxelVolume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, featL, featR, left, right, p2, p3, fb, trans, trans_inv, input_h, input_w):
        lib = _lib.load()
        featL, featR = _chk(featL, "featL"), _chk(featR, "featR")
        geo = [_chk(t, n) for t, n in ((left, "left_boxes"), (right, "right_boxes"), (p2, "p2"), (p3, "p3"), (fb, "fb"),
                                        (trans, "trans"), (trans_inv, "trans_inv"))]
        B, C, H, W = featL.shape
        N = geo[0].shape[0]
        voxel = torch.empty((N, 3 * C, 10, 10, 10), device=featL.device, dtype=_F32)
        dori = torch.empty((N,), device=featL.device, dtype=_F32)
        nws = lib.side_voxel_volume_ws_bytes(B, C, H, W)
        ws = torch.empty((max(nws, 16),), device=featL.device, dtype=torch.uint8)
        flags = _lib.VOXEL_ALIGN_CORNERS if voxel_align_corners else 0
        _lib.check(lib.side_voxel_volume_fwd(featL.data_ptr(), featR.data_ptr(), *[g.data_ptr() for g in geo], voxel.data_ptr(),
                                             dori.data_ptr(), N, B, C, H, W, int(input_h), int(input_w), flags, ws.data_ptr(), nws,
                                             _stream()), "side_voxel_volume_fwd")
        ctx.save_for_backward(*geo)
        ctx.meta = (B, C, H, W, int(input_h), int(input_w), flags)
        ctx.mark_non_differentiable(dori)
        return voxel, dori

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gvoxel, _gd):
        lib = _lib.load()
        geo = ctx.saved_tensors
        B, C, H, W, ih, iw, flags = ctx.meta
        gvoxel = _chk(gvoxel, "grad_voxel")
        N = geo[0].shape[0]
        gL = torch.empty((B, C, H, W), device=gvoxel.device, dtype=_F32)
        gR = torch.empty((B, C, H, W), device=gvoxel.device, dtype=_F32)
        nws = lib.side_voxel_volume_ws_bytes(B, C, H, W)
        ws = torch.empty((max(nws, 16),), device=gvoxel.device, dtype=torch.uint8)
        _lib.check(lib.side_voxel_volume_bwd(gvoxel.data_ptr(), *[g.data_ptr() for g in geo], gL.data_ptr(), gR.data_ptr(), N, B, C,
                                             H, W, ih, iw, flags, ws.data_ptr(), nws, _stream()), "side_voxel_volume_bwd")
        return (gL, gR) + (None,) * 9


def voxel_volume(featL, featR, left, right, p2, p3, fb, trans, trans_inv, input_h=384, input_w=1280):
    """Fused get_voxel + grid_sample + mask + cat of stereo_network_new.forward (:409-449):
    feats [B,64,H,W], boxes [N,5] -> (voxel [N,192,10,10,10] = cat(L - R, L, R), depth_ori [N])."""
    f32 = lambda t: t.to(_F32).contiguous()
    return _VoxelVolume.apply(featL, featR, f32(left), f32(right), f32(p2), f32(p3), f32(fb.reshape(-1)), f32(trans), f32(trans_inv),
                              input_h, input_w)


def conv_wgrad_tc(x_hi, x_lo, gy, Cin, ksize, stride=1):
    """Weight gradient of a convolution from the fp16 (hi, lo) channels-last input pairs the forward used and grad_output
    [N, Cout, Do, Ho, Wo] (side_conv_wgrad_tc): -> [Cout, Cin, kd, kh, kw]."""
    lib = _lib.load()
    x_hi, x_lo = _chk(x_hi, "x_hi", torch.float16), _chk(x_lo, "x_lo", torch.float16)
    gy = _chk(gy, "grad_output")
    N, D, H, W, Cp = x_hi.shape
    Cout = gy.shape[1]
    kd, kh, kw = ksize
    taps = kd * kh * kw
    nws = lib.side_conv_wgrad_tc_ws_bytes(N, D, H, W, Cp, Cout, kd, kh, kw, int(stride))
    if nws == 0:
        raise RuntimeError("side_conv_wgrad_tc failed (code -5): unsupported shape")
    ws = torch.empty((nws + 256,), device=gy.device, dtype=torch.uint8)
    off = (-ws.data_ptr()) % 256
    gw = torch.empty((Cout, taps, (Cp + 63) // 64 * 64), device=gy.device, dtype=_F32)
    _lib.check(lib.side_conv_wgrad_tc(x_hi.data_ptr(), x_lo.data_ptr(), gy.data_ptr(), gw.data_ptr(), N, D, H, W, Cp, Cout, kd, kh, kw,
                                      int(stride), ws.data_ptr() + off, nws, _stream()), "side_conv_wgrad_tc")
    return gw[:, :, :Cin].permute(0, 2, 1).reshape(Cout, Cin, kd, kh, kw).contiguous()
