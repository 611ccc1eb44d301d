/*
 * side_b200.h -- C ABI of libside_b200.so: the B200 (sm_100a) implementation of SIDE's stereo hot path.
 *
 * This is the drop-in boundary.  The reference's only native boundary is the pybind11 module `_ext`
 * (DCNv2/src/vision.cpp:4-9) with dcn_v2_forward / dcn_v2_backward (DCNv2/src/dcn_v2.h:9-73); everything
 * else on the path is Python calling ATen/torchvision ops.  Each entry point below names the reference
 * interface it replaces.  No torch / ATen types appear here: plain device pointers, ints and a stream.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to C-contiguous float32 (or the stated integer type) unless
 *     marked "host"; tensors are NCHW as in the reference;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is enqueued
 *     on it, nothing synchronises the host, the library is re-entrant (no global mutable state except
 *     one-time cudaFuncSetAttribute calls) -- see SURVEY.md section 8b "Threading / streams";
 *   - return value: 0 (SIDE_OK) or a negative SIDE_ERR_* code; side_last_error() gives a
 *     thread-local message.  The Python layer maps non-zero to RuntimeError like AT_ASSERTM/AT_ERROR
 *     do in the reference (dcn_v2_cuda.cu:61-85);
 *   - there is NO CPU implementation: host pointers are rejected (SIDE_ERR_NOT_DEVICE), mirroring
 *     AT_ERROR("Not implemented on the CPU") (dcn_v2.h:38).
 */
#ifndef SIDE_B200_H_
#define SIDE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIDE_ABI_VERSION 1

enum {
    SIDE_OK = 0,
    SIDE_ERR_INVALID_ARG = -1,   /* bad shape / null pointer / unsupported parameter combination */
    SIDE_ERR_NOT_DEVICE = -2,    /* a tensor pointer is not device memory */
    SIDE_ERR_WORKSPACE = -3,     /* workspace missing or too small */
    SIDE_ERR_CUDA = -4,          /* a CUDA runtime call or kernel launch failed */
    SIDE_ERR_UNSUPPORTED = -5    /* valid request this build cannot serve (e.g. tcgen05 path for Cin % 32 != 0) */
};

/* flags for side_dcn_fwd / side_dcn_bwd */
enum {
    SIDE_DCN_MASK_IS_LOGIT = 1 << 0, /* `mask` holds pre-sigmoid logits: sigmoid is fused (DCN.forward, dcn_v2.py:122) */
    SIDE_DCN_FUSE_AFFINE   = 1 << 1, /* y = y*scale[o] + shift[o]  (eval-mode BatchNorm of DeformConv.actf,
                                        feature_extraction_dla34.py:348-351) */
    SIDE_DCN_FUSE_RELU     = 1 << 2, /* y = max(y, 0) after the affine */
    SIDE_DCN_PREC_FP32     = 0 << 4, /* SIMT fp32 FMA implicit GEMM (<=1e-5 rel vs the reference) */
    SIDE_DCN_PREC_3XTF32   = 1 << 4, /* tcgen05 kind::tf32, hi/lo split, 3 MMAs: fp32-class accuracy (<=1e-4 rel); falls back to
                                        the SIMT kernel for shapes tcgen05 cannot tile.  What callers should pass. */
    SIDE_DCN_PREC_TF32     = 2 << 4, /* tcgen05 kind::tf32 single pass (~1e-3 rel, opt-in) */
    SIDE_DCN_PREC_3XFP16   = 3 << 4, /* tcgen05 kind::f16 on fp16 (hi, lo * 2^11) operand pairs, 3 MMAs at twice the tf32 rate: fp32-class
                                        accuracy (<=1e-4 rel) while |x|, |w| stay in fp16's range -- reported through
                                        side_tc_range_guard; Cin % 64 != 0 runs as 3xTF32 */
    SIDE_DCN_PREC_MASK     = 3 << 4,
    SIDE_DCN_BWD_SIMT_GEMM = 1 << 9, /* side_dcn_bwd: keep the column GEMM of the channels-last path on the fp32 SIMT kernel
                                        (default: tcgen05 3xTF32 when Cout % 32 == 0 and the pixel count tiles by 128) */
    SIDE_DCN_BWD_SCALAR    = 1 << 8  /* side_dcn_bwd: force the scalar-atomic kernel that keeps the reference's thread mapping
                                        (one thread per (pixel, tap), serial over channels); default is the channels-last
                                        path with 16-byte vector reductions whenever dg == 1, Cin % 64 == 0, P % 4 == 0 */
};

/* flags for side_inst_costvol_fwd / _bwd */
enum {
    SIDE_VOL_GATE = 1 << 0, /* multiply every (roi, depth) slice by the cosine gate x_cross
                               (cost_volume.forward, stereo_network_old.py:197-203) */
    SIDE_VOL_FMA = 1 << 1,  /* opt-in: contract the 4-tap bilinear sum into FMAs (as nvcc does for torchvision's CUDA
                               kernel).  <= 1e-6 relative to the bit-exact default, ~half the instructions. */
    SIDE_VOL_SEPARABLE = 1 << 2, /* separable evaluation: the y interpolation (identical for all D candidates of a RoI)
                               is computed once per RoI into shared memory, every bin is then 4 taps.  Values differ
                               from torchvision's operation order by a few ulp (<= 1e-5 relative);
                               channel placement and L-R stay exact.  Needs P == 16, C % 8 == 0, D <= 256 and
                               side_inst_costvol_fast_ws_bytes(...) bytes of workspace (the per-CTA gate-statistics partials;
                               the features are read in place as NCHW, 16-byte loads along x when W % 4 == 0). */
    SIDE_VOL_XCROSS = 1 << 3, /* with SEPARABLE and without GATE: also return the gate scalar xcross[N, D] computed in the
                               same pass, WITHOUT applying it (the consumer, side_ncdhw_to_cl_split, multiplies) */
    SIDE_VOL_BWD_SCALAR = 1 << 4, /* side_inst_costvol_bwd: force the scalar-atomic kernel (torchvision's roi_align backward
                               thread mapping, 32 atomics per volume element); default for P == 16, C % 8 == 0 is the
                               separable gather kernel that keeps the x-pass sums in registers */
    SIDE_VOL_FEAT_NHWC = 1 << 5, /* side_inst_costvol_fwd_cl: featL / featR are already channels-last [B, H, W, C] (e.g. the output of
                               side_conv3d_tc_fwd): the two staging copies are skipped */
    SIDE_VOL_NO_DIFF = 1 << 6  /* side_inst_costvol_fwd_cl: emit only the L and R planes, [N, D, P, P, 2C].  The third plane of
                               cat(L, R, L - R) (stereo_network_old.py:374-376) is linear in the other two; the volume's only
                               consumer, the first convolution of cost_volume.dres0 (:147-149), then runs with the folded
                               weights (W_L + W_D, W_R - W_D) on 2C input channels: same function, <= 1e-6 of the output range */
};

/* decode flavour */
enum {
    SIDE_DECODE_HEAT_IS_LOGIT = 1 << 0 /* apply sigmoid to the heat-map first (bbox_decode, decode.py:93) */
};

int side_abi_version(void);
const char *side_last_error(void);
/* 1 when the shared object carries sm_100a device code and device 0 can run it; never fails. */
int side_device_ok(void);

/* ---------------------------------------------------------------------------------------------
 * DCNv2 modulated deformable convolution.
 * Replaces _ext.dcn_v2_forward (DCNv2/src/dcn_v2.h:9-39 -> dcn_v2_cuda_forward, dcn_v2_cuda.cu:43-173
 * + modulated_deformable_im2col_gpu_kernel, dcn_v2_im2col_cuda.cu:125-195).  No `columns` buffer is
 * materialised: the modulated bilinear gather feeds the contraction tile by tile.
 *   x      [B, Cin, H, W]
 *   offset [B, dg*2*kh*kw, Ho, Wo]  batch stride offset_bs floats (0 = dense); channel 2k = dy, 2k+1 = dx
 *   mask   [B, dg*kh*kw,   Ho, Wo]  batch stride mask_bs floats (0 = dense); logits if MASK_IS_LOGIT
 *          (offset / mask may alias the 27-channel conv_offset_mask output: offset = om, mask = om + 18*Ho*Wo,
 *           both with batch stride 27*Ho*Wo -- the chunk/cat of dcn_v2.py:120-121 is a no-op on memory)
 *   w      [Cout, Cin, kh, kw], bias [Cout] (may be NULL)
 *   scale/shift [Cout] used only with SIDE_DCN_FUSE_AFFINE
 *   y      [B, Cout, Ho, Wo]
 *   ws     workspace of side_dcn_fwd_ws_bytes(...) bytes (re-laid-out weights); may be NULL when that is 0.
 * Supported: any kh/kw/stride/pad/dilation; deformable_groups dg > 1 when (Cin / dg) % 16 == 0 (fp32 SIMT kernel).  The tcgen05
 * precisions need dg == 1, Cin % 32 == 0, Cout % 16 == 0, Cout <= 256; for other shapes they run the fp32 SIMT kernel, so
 * SIDE_DCN_PREC_3XTF32 is safe to pass always (it is what side_b200.ops passes by default).
 * --------------------------------------------------------------------------------------------- */
size_t side_dcn_fwd_ws_bytes(int B, int Cin, int H, int W, int Cout, int kh, int kw, int flags);
int side_dcn_fwd(const float *x, const float *offset, const float *mask, const float *w, const float *bias,
                 const float *scale, const float *shift, float *y, int B, int Cin, int H, int W, int Cout, int kh,
                 int kw, int sh, int sw, int ph, int pw, int dh, int dw, int dg, long long offset_bs,
                 long long mask_bs, int flags, void *ws, size_t ws_bytes, void *stream);

/* Channels-last front end of the tcgen05 path (inference): x_nhwc [B, H, W, Cin] is already channels-last and
 * om_cl [B, Ho, Wo, om_ld] is the channels-last output of DCN.conv_offset_mask (dcn_v2.py:105-123; channels 0..2*kh*kw-1 =
 * interleaved (dy, dx), then kh*kw mask logits), e.g. computed by side_conv3d_tc_fwd with Cout padded to 32.  y is NCHW.
 * Only SIDE_DCN_PREC_3XTF32 / _TF32; flags may add FUSE_AFFINE / FUSE_RELU (MASK_IS_LOGIT is implied). */
int side_dcn_fwd_cl(const float *x_nhwc, const float *om_cl, int om_ld, const float *w, const float *bias, const float *scale,
                    const float *shift, float *y, int B, int Cin, int H, int W, int Cout, int kh, int kw, int sh, int sw, int ph,
                    int pw, int dh, int dw, int flags, void *ws, size_t ws_bytes, void *stream);

/* Replaces _ext.dcn_v2_backward (dcn_v2.h:41-73 -> dcn_v2_cuda_backward, dcn_v2_cuda.cu:207-336 and the
 * col2im / col2im_coord kernels, dcn_v2_im2col_cuda.cu:197-327).  All five gradients are OVERWRITTEN
 * (the reference zero-initialises them, :252-256).  grad_mask is w.r.t. the post-sigmoid mask unless
 * MASK_IS_LOGIT is set, in which case it is w.r.t. the logits.  Any output pointer may be NULL to skip it. */
size_t side_dcn_bwd_ws_bytes(int B, int Cin, int H, int W, int Cout, int kh, int kw, int flags);
int side_dcn_bwd(const float *x, const float *offset, const float *mask, const float *w, const float *gy,
                 float *gx, float *goffset, float *gmask, float *gw, float *gbias, int B, int Cin, int H, int W,
                 int Cout, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int dg,
                 long long offset_bs, long long mask_bs, int flags, void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Depth-candidate / shifted-RoI generation.  Replaces get_proposal_shift (stereo_network_old.py:34-133).
 *   left, right [N,5] = (b, x1, y1, x2, y2), grouped by image in ascending b;  fb [B]
 *   pro_left, pro_right [D, N, 5], depth_bin [N, D];  x_clamp = input_w//4 - 1 (319 for 1280-wide input)
 * --------------------------------------------------------------------------------------------- */
int side_proposal_shift(const float *left, const float *right, const float *fb, int N, int B, int D,
                        float x_clamp, float *pro_left, float *pro_right, float *depth_bin, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Instance-level depth cost volume with RoIAlign sampling fused in.
 * Replaces the loop stereo_network_old.py:368-376 (2*D torchvision RoIAlign launches + 3*D slice copies)
 * and, with SIDE_VOL_GATE, the cosine gate of cost_volume.forward (:197-203).
 *   featL, featR [B, C, H, W];  left, right [N,5] boxes (as for side_proposal_shift);  fb [B]
 *   valid        [N] uint8 or NULL: rows with valid==0 produce an all-zero volume slice, depth_bin = 0
 *   cost         [N, 3C, D, P, P];  depth_bin [N, D];  xcross [N, D] (NULL allowed; written only with GATE)
 * RoIAlign semantics: torchvision legacy aligned=False, spatial_scale 1, sampling_ratio 2 (:271).
 *   ws           optional workspace of side_inst_costvol_ws_bytes(B,C,H,W) bytes: enables the channels-last gather
 *                (the features are transposed once to NHWC so each bilinear tap is one 16-byte load for 4 channels);
 *                NULL selects the slower NCHW gather.  Results are bit-identical either way.
 * --------------------------------------------------------------------------------------------- */
size_t side_inst_costvol_ws_bytes(int B, int C, int H, int W);
size_t side_inst_costvol_fast_ws_bytes(int B, int C, int H, int W, int N, int D);
int side_inst_costvol_fwd(const float *featL, const float *featR, const float *left, const float *right,
                          const float *fb, const uint8_t *valid, float *cost, float *depth_bin, float *xcross,
                          int N, int B, int C, int H, int W, int D, int P, float x_clamp, int flags, void *ws,
                          size_t ws_bytes, void *stream);

/* Backward of the above w.r.t. featL / featR (autograd of RoIAlign + CopySlices + gate in the reference).
 * gfeatL / gfeatR [B,C,H,W] are ACCUMULATED into (caller zero-fills); fp32 atomics => summation order is not
 * deterministic, as in torchvision's roi_align backward. */
int side_inst_costvol_bwd(const float *featL, const float *featR, const float *left, const float *right,
                          const float *fb, const uint8_t *valid, const float *gcost, float *gfeatL, float *gfeatR,
                          int N, int B, int C, int H, int W, int D, int P, float x_clamp, int flags, void *stream);

/* Same gradients with a workspace: for SIDE_VOL_GATE the raw volume is recomputed by the separable forward, the gate's
 * backward runs in place on it and the separable gather kernel (register sums, one atomic per window pixel) finishes --
 * several times faster than the scalar kernel above.  Falls back to side_inst_costvol_bwd when P != 16 or C % 8 != 0.
 * ws_bytes >= side_inst_costvol_bwd_fast_ws_bytes(...) (0 without the gate). */
size_t side_inst_costvol_bwd_fast_ws_bytes(int B, int C, int H, int W, int N, int D, int flags);
int side_inst_costvol_bwd_fast(const float *featL, const float *featR, const float *left, const float *right,
                               const float *fb, const uint8_t *valid, const float *gcost, float *gfeatL, float *gfeatR,
                               int N, int B, int C, int H, int W, int D, int P, float x_clamp, int flags, void *ws,
                               size_t ws_bytes, void *stream);

/* The same volume emitted directly in the format its consumer (side_conv3d_tc_fwd_f16, cost_volume.dres0) reads: channels-last
 * [N, D, P, P, 3C] and already split into fp16 operand pairs, with the cosine gate applied when SIDE_VOL_GATE is set -- one pass
 * over HBM instead of volume write + re-read + re-write (stereo_network_old.py:366-376 and :197-203 in one kernel).
 *   cost_hi, cost_lo  half [N, D, P, P, 3C]: hi = fp16(v), lo = fp16((v - hi) * 2^11), channels [L (C) | R (C) | L - R (C)]
 *   depth_bin [N, D];  xcross [N, D] (may be NULL): the gate scalar of every slice, whether or not it was applied
 * Built for C == 32 (the reference's reduced_channel), P == 16, 2 <= D <= 64; other shapes return SIDE_ERR_INVALID_ARG (callers use
 * side_inst_costvol_fwd + side_ncdhw_to_cl_split_f16).  Values follow the separable evaluation order (<= 1e-5 relative to
 * torchvision's RoIAlign; L - R from the kernel's own L and R).  ws: side_inst_costvol_cl_ws_bytes(...) bytes, 16-byte aligned. */
size_t side_inst_costvol_cl_ws_bytes(int B, int C, int H, int W);
int side_inst_costvol_fwd_cl(const float *featL, const float *featR, const float *left, const float *right, const float *fb,
                             const uint8_t *valid, void *cost_hi, void *cost_lo, float *depth_bin, float *xcross, int N, int B,
                             int C, int H, int W, int D, int P, float x_clamp, int flags, void *ws, size_t ws_bytes, void *stream);

/* Stand-alone cosine gate on an already built volume (drop-in cost_volume.forward(cost, ...) entry,
 * stereo_network_old.py:194-203).  out may alias cost.  bwd: gcost = d loss / d cost. */
int side_xcross_gate_fwd(const float *cost, float *out, float *xcross, int N, int C, int D, int P, void *stream);
int side_xcross_gate_bwd(const float *cost, const float *gout, float *gcost, int N, int C, int D, int P,
                         void *stream);

/* ---------------------------------------------------------------------------------------------
 * Soft-argmin tail.  Replaces stereo_network_old.py:228-236 (AvgPool2d(S) -> softmax over D -> sum p*depth_bin).
 *   logits [N, D, S, S] (squeezed classify output, S = 4);  depth_bin [N, D];  depth [N];  prob [N, D] (NULL ok)
 * --------------------------------------------------------------------------------------------- */
int side_softargmin_fwd(const float *logits, const float *depth_bin, float *depth, float *prob, int N, int D,
                        int S, void *stream);
int side_softargmin_bwd(const float *prob, const float *depth_bin, const float *depth, const float *gdepth,
                        float *glogits, float *gdepth_bin, int N, int D, int S, void *stream);

/* ---------------------------------------------------------------------------------------------
 * CenterNet decode: 3x3 max-pool NMS + per-class top-K + cross-class top-K + head gathers, ONE launch.
 * Replaces _nms/_topk/_gather_feat/_transpose_and_gather_feat (decode.py:9-33, utils.py:12-26) and the
 * box / detection assembly of bbox_decode (decode.py:91-126) and ddd_decode (:35-89).
 * Tie rule: larger score first, then lower flat index (torch.topk leaves ties unspecified).
 *   heat [B, Cat, H, W]; K <= 1024; Cat*K <= 8192
 *   ws: side_decode_ws_bytes(B, Cat, K) bytes of scratch (zero-initialised by the call itself)
 * bbox flavour:  wh, reg [B,3,H,W] -> bbox, bbox_right [B,K,5] = (b, x1,y1,x2,y2); keep [B*K] uint8 = (sum of the
 *                4 coords > 0, decode.py:123); slot [B*K] int32 = rank of the row among its image's kept rows;
 *                count [B] int32 = kept rows per image.  wh is multiplied by wh_scale (stereo_network_old.py:360).
 * ddd flavour:   kept [B,6*grid,H,W], dim [B,3,..], orien [B,2,..], wh, reg -> det, det_right [B,K,6], info [B,K,9]
 *                info[...,8] = floor(argmax / grid).
 * Both also return score [B,K], ind [B,K] int32 (flat y*W+x), cls [B,K] int32 (any may be NULL).
 * --------------------------------------------------------------------------------------------- */
size_t side_decode_ws_bytes(int B, int Cat, int K);
int side_bbox_decode(const float *heat, const float *wh, const float *reg, float *bbox, float *bbox_right,
                     uint8_t *keep, int32_t *slot, int32_t *count, float *score, int32_t *ind, int32_t *cls, int B,
                     int Cat, int H, int W, int K, float wh_scale, int flags, void *ws, size_t ws_bytes,
                     void *stream);
int side_ddd_decode(const float *heat, const float *kept, const float *dim, const float *orien, const float *wh,
                    const float *reg, float *det, float *det_right, float *info, float *score, int32_t *ind,
                    int32_t *cls, int B, int Cat, int H, int W, int grid, int K, int flags, void *ws,
                    size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Full-image disparity-sweep volumes (north_star item 1; no call site in the reference, SURVEY.md F3/A6;
 * the reference's stereo-head concat torch.cat((L,R),1), stereo_network_old.py:348, is the D=1 case).
 *   concat [B, 2C, D, H, W]: [:, 0:C, d, y, x] = L[.., x]*[x>=d];  [:, C:2C, d, y, x] = R[.., x-d]*[x>=d]
 *   gwc    [B, G,  D, H, W]: mean over the C/G channels of group g of L[.., x]*R[.., x-d], 0 where x<d
 * Backward entry points OVERWRITE gL, gR [B,C,H,W] (deterministic gather formulation, no atomics).
 * --------------------------------------------------------------------------------------------- */
int side_concat_volume_fwd(const float *L, const float *R, float *vol, int B, int C, int H, int W, int D,
                           void *stream);
int side_concat_volume_bwd(const float *gvol, float *gL, float *gR, int B, int C, int H, int W, int D,
                           void *stream);
int side_gwc_volume_fwd(const float *L, const float *R, float *vol, int B, int C, int H, int W, int D, int G,
                        void *stream);
int side_gwc_volume_bwd(const float *L, const float *R, const float *gvol, float *gL, float *gR, int B, int C,
                        int H, int W, int D, int G, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Depth-wise transposed convolution of the DLA up-sampling neck.
 * Replaces IDAUp.up_k = nn.ConvTranspose2d(o, o, 2f, stride=f, padding=f//2, groups=o, bias=False)
 * (feature_extraction_dla34.py:370-373), which cuDNN serves with a generic grouped direct kernel.
 *   x [B,C,H,W], w [C,1,k,k] -> y [B,C,Ho,Wo], Ho = (H-1)*stride - 2*pad + k
 * bwd OVERWRITES gx [B,C,H,W] and gw [C,1,k,k] (either may be NULL).
 * --------------------------------------------------------------------------------------------- */
int side_dw_deconv_fwd(const float *x, const float *w, float *y, int B, int C, int H, int W, int k, int stride, int pad,
                       void *stream);
int side_dw_deconv_bwd(const float *x, const float *w, const float *gy, float *gx, float *gw, int B, int C, int H, int W,
                       int k, int stride, int pad, void *stream);

/* One IDAUp step between its two deformable convolutions, fused (inference; feature_extraction_dla34.py:380-386):
 *     sum = up_k(x) + skip            up_k = the depth-wise ConvTranspose2d above with k = 2 * stride, pad = stride / 2
 * written channels-last [B, H*stride, W*stride, Cpad] three times over: `full` fp32 (what node_k's deformable gather reads,
 * side_dcn_fwd_cl) and the fp16 operand pairs hi / lo (what node_k's offset convolution reads, side_conv3d_tc_fwd_f16).
 * Replaces side_dw_deconv_fwd + the elementwise addition + side_ncdhw_to_cl_split_f16; bit-identical to that sequence.
 *   x [B,C,H,W] (output of proj_k), w [C,1,2*stride,2*stride], skip [B,C,H*stride,W*stride]; stride in {2, 4, 8};
 *   Cpad >= C, even: channels C..Cpad-1 are written as zeros (Cpad = C rounded up to 32 in the network). */
int side_idaup_fuse_cl_f16(const float *x, const float *w, const float *skip, float *full, void *hi, void *lo, int B, int C,
                           int H, int W, int stride, int Cpad, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Aggregation network of the instance-depth branch on the tensor cores.
 * Replaces the cuDNN calls behind cost_volume.dres0 / dres1 / dres2 / classify (stereo_network_old.py:139-171,
 * 205-227): nn.Conv3d(k=3, stride 1, padding 1, bias=False) [+ BatchNorm3d (eval) + ReLU (+ residual)].
 * Activations are channels-last (NDHWC) and kept as the exact tf32 split x = hi + lo (hi = top 19 bits) so that
 * the 3xTF32 contraction (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM) matches fp32 to <= 1e-4 relative.
 *   side_conv_tc_prep_weights: w [Cout, Cin, taps] (= nn.Conv3d.weight viewed [Cout, Cin, 27]) -> wp, the per-k-block
 *       swizzled shared-memory image (hi and lo), side_conv_tc_weight_bytes(Cin, Cout, taps) bytes.  Once per layer.
 *   side_conv3d_tc_fwd: x_hi, x_lo [N, D, H, W, Cin] -> y [N, D, H, W, Cout] (fp32, may be NULL) and / or its split
 *       y_hi, y_lo (may be NULL); relu = 1: out = relu(conv * scale[o] + shift[o]) + residual (dres2(cost) + cost, :216);
 *       relu = 2: out = relu(conv * scale[o] + shift[o] + residual) (DLA BasicBlock); relu = 0: no activation.
 *       relu | 4: MaxPool3d((1,2,2)) (cost_volume.max_pool1 / max_pool2, :213, :218) of that result fused into the epilogue;
 *       all outputs are then [N, D, H/2, W/2, Cout] (residual stays at the convolution's resolution).  Stride 1, boxes of
 *       2..16 columns x an even number of rows (16x16 and 8x8 maps).
 *       stride_hw = 2 (2-D kernels only): [N, D, H/2, W/2, Cout] outputs; the TMA box then traverses the input with
 *       element strides, so a strided convolution costs no more than a dense one.
 *       Needs Cin % 32 == 0; Cout % 16 == 0 (<= 128) or Cout % 128 == 0 (<= 1536, processed as 128-wide n-tiles); a
 *       (D, H, W) that tiles into 128-voxel boxes (box w = largest power of two <= 128 dividing W, then rows, then
 *       slices: 16x16, 8x8 with even D, 4x4 with D % 8 == 0, 96x320 as 2 rows x 64 columns; for 2-D kernels, where D is the
 *       batch, the last box may hang over the end of the batch: 2 images at the 12x40 DLA level use boxes of 4); kernel 3x3x3, 1x3x3 or
 *       1x1x1 (2-D convolutions are kd = 1 with the batch as D or D = 1: the head convolutions of
 *       stereo_network.forward, :343-348, and the DLA-34 levels 2-5 of feature_extraction_dla34.py).
 * Helpers (one pass over HBM each):
 *   side_ncdhw_to_cl_split  x [N, C, S] (* scale[N, D], D | S, or NULL) -> hi, lo [N, S, C]   (volume from
 *                           side_inst_costvol_fwd; scale = xcross applies the gate deferred by SIDE_VOL_XCROSS)
 *   side_tf32_split         x -> hi, lo, n elements (n % 4 == 0)
 *   side_gate_mul_split     y [N,D,H,W,C] * gate [N,D,W,C] -> hi, lo              (isp * cost, :207-210)
 *   side_maxpool_hw2_cl     x [N,D,H,W,C] -> MaxPool3d((1,2,2)) -> y and / or hi, lo  (:213, :218)
 *   side_conv3d_c1_cl       x [N,D,H,W,C], w [1,C,3,3,3] -> out [N,D,H,W]         (classify.3, :170)
 * --------------------------------------------------------------------------------------------- */
size_t side_conv_tc_weight_bytes(int Cin, int Cout, int taps);
int side_conv_tc_prep_weights(const float *w, float *wp, int Cout, int Cin, int taps, void *stream);
int side_conv3d_tc_fwd(const float *x_hi, const float *x_lo, const float *wp, const float *scale, const float *shift,
                       const float *residual, float *y, float *y_hi, float *y_lo, int N, int D, int H, int W, int Cin,
                       int Cout, int kd, int kh, int kw, int stride_hw, int relu, void *stream);
/* test / benchmark hook, bit mask: 0 = default; 1 = halo-box reuse in the voxel-major kernel; 32 = disable the role-swapped
 * Cout = 64 kernel.  Results are identical up to fp32 summation order. */
int side_conv_tc_set_mode(int mode);
int side_ncdhw_to_cl_split(const float *x, const float *scale, float *full, float *hi, float *lo, int N, int C, long long S,
                           int D, void *stream);   /* full (may be NULL): the unsplit channels-last copy as well */
int side_tf32_split(const float *x, float *hi, float *lo, long long n, void *stream);
int side_f16_split(const float *x, void *hi, void *lo, long long n, void *stream);   /* fp16 (hi, lo * 2^11) pairs, same layout */
int side_gate_mul_split(const float *y, const float *gate, float *hi, float *lo, int N, int D, int H, int W, int C,
                        void *stream);
int side_maxpool_hw2_cl(const float *x, float *y, float *hi, float *lo, int N, int D, int H, int W, int C, void *stream);

/* "3xFP16" variants of the five entries above: the operand pairs are fp16 arrays (hi = fp16(x), lo = fp16((x - hi) * 2^11)) and
 * the MMAs are kind::f16 -- the same 22 significand bits per operand as the tf32 pair, fp32 accumulation, hi*hi in the main
 * accumulator and the two cross terms in a second one that the epilogue scales by 2^-11; twice the tensor rate of 3xTF32 and
 * half the operand bytes.  Needs Cin % 32 == 0 (k-blocks are 64 channels: prepare the weights with Cin zero-padded to a
 * multiple of 64; a half-full last block is zero-filled by TMA and its empty MMAs are skipped) and |x| < 65504
 * (saturating conversions; activations behind a BatchNorm).
 * Weight tiles take side_conv_tc_weight_bytes(...) / 2 bytes.  y, residual, scale, shift stay fp32. */
int side_conv_tc_prep_weights_f16(const float *w, void *wp, int Cout, int Cin, int taps, void *stream);
int side_conv3d_tc_fwd_f16(const void *x_hi, const void *x_lo, const void *wp, const float *scale, const float *shift,
                           const float *residual, float *y, void *y_hi, void *y_lo, int N, int D, int H, int W, int Cin, int Cout,
                           int kd, int kh, int kw, int stride_hw, int relu, void *stream);
int side_ncdhw_to_cl_split_f16(const float *x, const float *scale, float *full, void *hi, void *lo, int N, int C, long long S,
                               int D, int Cpad /* channels per output voxel >= C, the extra ones zero (k-blocks are 64 wide) */,
                               void *stream);
int side_gate_mul_split_f16(const float *y, const float *gate, void *hi, void *lo, int N, int D, int H, int W, int C, void *stream);
int side_maxpool_hw2_cl_f16(const float *x, float *y, void *hi, void *lo, int N, int D, int H, int W, int C, void *stream);
int side_conv3d_c1_cl(const float *x, const float *w, float *out, int N, int D, int H, int W, int C, void *stream);
/* Range guard of the 3xFP16 operand pairs.  fp16 covers 2^-24 .. 65504, the reference's fp32 far more.  Registers a ring of
 * `nwords` zero-initialised uint32 slots in device memory for the CURRENT device (NULL unregisters).  From then on every entry that
 * writes fp16 pairs (side_ncdhw_to_cl_split_f16, side_gate_mul_split_f16, side_maxpool_hw2_cl_f16, side_conv3d_tc_fwd_f16 with
 * y_hi/y_lo, side_inst_costvol_fwd_cl) takes the next slot (round robin) and leaves there the float bit pattern of max |x| of
 * the tensor it split (atomicMax: slots only grow).  The caller classifies the ring when it chooses -- a slot >= 0x477FE000
 * (65504.0f): a saturating conversion happened, results invalid; a non-zero slot < 0x36800000 (2^-18): the 2^-36 absolute
 * resolution of the pair no longer meets the 1e-4 parity bar -- zeroes it, and on either finding repeats the work with the
 * 3xTF32 entries (8-bit exponent).  With more than `nwords` guarded launches between two readings slots are shared (maxima
 * merge: saturation is still caught, an underflow may be masked). */
int side_tc_range_guard(void *status_words, int nwords);
/* channels-last [B, HW, C] -> NCHW [B, C, HW]: hands a tensor-core convolution output back to NCHW consumers */
int side_cl_to_nchw(const float *x, float *y, int B, int C, long long HW, void *stream);
/* channel concatenation of channels-last tensors: srcs[i] = [rows, row_bytes[i]] (host arrays of device pointers / byte counts, 1..8
 * sources, rows multiples of 16 bytes) -> dst [rows, sum row_bytes]; 16-byte copies (Root inputs of the DLA trees) */
int side_cl_concat(const void *const *srcs, const int *row_bytes, int nsrc, void *dst, long long rows, void *stream);
/* same from rows of ld >= C floats: the first C channels of every row (outputs computed with padded channel counts) */
int side_cl_to_nchw_ld(const float *x, int ld, float *y, int B, int C, long long HW, void *stream);

/* ---------------------------------------------------------------------------------------------
 * DLA-34 stem: Conv2d(k, stride, padding (k-1)/2, bias=False) + eval-mode BatchNorm2d
 * (folded: scale, shift; NULL = identity) + ReLU as one direct fp32 convolution, NCHW in / out.  Built for the three
 * stem layers of feature_extraction_dla34.py (DLA.base_layer 3->16 k7 s1, level0 16->16 k3 s1, level1 16->32 k3 s2);
 * other shapes return SIDE_ERR_UNSUPPORTED (the caller keeps cuDNN).
 *   x [B, Cin, H, W], w [Cout, Cin, k, k] -> y [B, Cout, Ho, Wo]
 * --------------------------------------------------------------------------------------------- */
int side_stem_conv_fwd(const float *x, const float *w, const float *scale, const float *shift, float *y, int B, int Cin, int H,
                       int W, int Cout, int k, int stride, int relu, void *stream);
/* base_layer only (3 -> 16, 7x7): the result as fp16 (hi, lo * 2^11) pairs in the 2x2 space-to-depth channels-last layout
 * [B, H/2, W/2, 64], channel = (dy * 2 + dx) * 16 + o, so that level0 (16 -> 16, 3x3) and level1 (16 -> 32, 3x3, stride 2) run as
 * 3x3 BLOCK convolutions (64 -> 64 and 64 -> 32, weights rearranged by the caller) on side_conv3d_tc_fwd_f16.  H, W even. */
int side_stem_conv_fwd_s2d(const float *x, const float *w, const float *scale, const float *shift, void *y_hi, void *y_lo,
                           int B, int H, int W, int relu, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Dense photometric alignment of the post-process path.
 * Replaces src/lib/dense_align/dense_align.py: the numpy normalisation + F.interpolate of align_parallel (:251-266),
 * sample() with its Box3d ray casting (:14-70, box_3d.py:9-102) and enumeration_depth (:175-237).
 * Images are PACKED texels: [H][W] float4 = (c0, c1, c2, 0), 16-byte aligned.
 *   side_dense_align_prep_u8   img_hwc uint8 [H,W,3] -> ((x/255) - mean)/std -> 2x bilinear (align_corners=False) -> packed [2H][2W]
 *                              mean3 / std3 are HOST pointers to 3 floats.
 *   side_dense_align_up2_pack  planar float [3,H,W] -> 2x bilinear -> packed [2H][2W]
 *   side_dense_align_pack / _unpack  planar [3,H,W] <-> packed [H][W]
 *   side_dense_align_sample    box_left [rois,4], borders [rois,2] (both already in the scaled image), poses [rois,7]
 *                              (x,y,z,w,h,l,theta), f / cx / cy of the scaled image, f_h x f_w its size ->
 *                              uvz [rois,cap,3] (u, v, depth offset of the visible box surface; valid pixels first, in the
 *                              reference's row-major order, zeros after), weight [rois,cap] (1 / 0), count [rois] (valid
 *                              pixels found; > cap means the tail was dropped).
 *   side_dense_align_enum      depth_enum [iters,rois], fb -> err_sum [iters,rois] = sum over pixels and channels of
 *                              |grid_sample(L) - grid_sample(R shifted by the hypothesis' disparity)| * weight,
 *                              best_depth [rois] = depth_enum[argmin (first minimum)], best_idx [rois] (may be NULL).
 *                              pixels = second dimension of uvz / weight.  flags: SIDE_DA_ALIGN_CORNERS selects
 *                              grid_sample(align_corners=True) (torch < 1.3 default, which the reference was written for);
 *                              0 = align_corners=False (what the reference computes under the torch in this image).
 * --------------------------------------------------------------------------------------------- */
#define SIDE_DA_ALIGN_CORNERS (1 << 0)
int side_dense_align_prep_u8(const unsigned char *img_hwc, float *out_packed, int H, int W, const float *mean3,
                             const float *std3, void *stream);
int side_dense_align_up2_pack(const float *img_chw, float *out_packed, int H, int W, void *stream);
int side_dense_align_pack(const float *img_chw, float *out_packed, int H, int W, void *stream);
int side_dense_align_unpack(const float *packed, float *img_chw, int H, int W, void *stream);
int side_dense_align_sample(const float *box_left, const float *borders, const float *poses, int rois, float f, float cx,
                            float cy, int f_h, int f_w, int cap, float *uvz, float *weight, int *count, void *stream);
int side_dense_align_enum(const float *imL_packed, const float *imR_packed, const float *uvz, const float *weight,
                          const float *depth_enum, float fb, int rois, int pixels, int iters, int H, int W, int flags,
                          float *err_sum, float *best_depth, int *best_idx, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Detector input preparation.  Replaces stereoDetector.pre_process's host work
 * (src/lib/modules/stereoDetector.py:45-82): cv2.warpAffine(INTER_LINEAR, constant border 0) of the raw uint8 image,
 * ((x / 255) - mean) / std and the HWC -> CHW transpose, for the left and right image in one launch.
 *   img_*      uint8 [src_h, src_w, 3] device pointers (img_right / out_right may be NULL)
 *   inv_map6   HOST pointer: the 2x3 map from DESTINATION to SOURCE pixels (cv2.invertAffineTransform(trans_input))
 *   mean3/std3 HOST pointers
 *   out_*      float [3, dst_h, dst_w]
 * OpenCV's 8-bit fixed-point algorithm is restated (1/32-pixel positions, 15-bit weights); cv2 is not in the image, so
 * this entry's parity is pinned only by the repo's own numpy restatement (oracle/torch_port.py:warp_affine_u8).
 * --------------------------------------------------------------------------------------------- */
int side_preprocess_u8(const unsigned char *img_left, const unsigned char *img_right, int src_h, int src_w,
                       const double *inv_map6, const float *mean3, const float *std3, float *out_left, float *out_right,
                       int dst_h, int dst_w, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Instance voxel volume of the stereo_network_new variant.
 * Replaces get_voxel (src/lib/models/networks/stereo_network_new.py:160-283, a host-side loop over images and RoIs) and
 * the per-image F.grid_sample / mask / cat sequence of stereo_network.forward (:409-449).
 *   left, right [N,5] boxes (b, x1, y1, x2, y2) in feature pixels; p2, p3 [B,3,4]; fb [B]; trans, trans_inv [B,2,3]
 *   input_h, input_w: the module constants the reference normalises with (u_max = input_w/4 - 1, v_max = input_h/4 - 1)
 *   side_voxel_coords      get_voxel's outputs: norm3 [N,10,10,10,3] + valid3 [N,10,10,10] (need depth_bins [N,D]; pass NULL for
 *                          all three to skip them), normL / normR [N,10,10,10,2], validL / validR [N,10,10,10], depth_ori [N]
 *   side_voxel_volume_fwd  featL, featR [B,64,H,W] -> voxel [N, 192, 10,10,10] = cat(L - R, L, R) of the bilinear samples
 *                          (padding zeros; voxels outside [-1,1] are zero), depth_ori [N] (may be NULL).
 *                          flags: SIDE_VOXEL_ALIGN_CORNERS = grid_sample(align_corners=True); 0 = False (torch >= 1.3 default,
 *                          what the reference computes today).  ws: side_voxel_volume_ws_bytes(B,C,H,W) bytes.
 *   side_voxel_volume_bwd  gvoxel -> gfeatL, gfeatR [B,64,H,W] (OVERWRITTEN; fp32 atomics inside).
 * --------------------------------------------------------------------------------------------- */
#define SIDE_VOXEL_ALIGN_CORNERS (1 << 0)
int side_voxel_coords(const float *left, const float *right, const float *p2, const float *p3, const float *fb,
                      const float *trans, const float *trans_inv, const float *depth_bins, int N, int B, int D, int H,
                      int W, int input_h, int input_w, float *norm3, float *valid3, float *normL, float *validL,
                      float *normR, float *validR, float *depth_ori, void *stream);
size_t side_voxel_volume_ws_bytes(int B, int C, int H, int W);
int side_voxel_volume_fwd(const float *featL, const float *featR, const float *left, const float *right, const float *p2,
                          const float *p3, const float *fb, const float *trans, const float *trans_inv, float *voxel,
                          float *depth_ori, int N, int B, int C, int H, int W, int input_h, int input_w, int flags,
                          void *ws, size_t ws_bytes, void *stream);
int side_voxel_volume_bwd(const float *gvoxel, const float *left, const float *right, const float *p2, const float *p3,
                          const float *fb, const float *trans, const float *trans_inv, float *gfeatL, float *gfeatR,
                          int N, int B, int C, int H, int W, int input_h, int input_w, int flags, void *ws,
                          size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Weight gradient of the plain convolutions on tcgen05 (training side of SURVEY.md 8f rows F1 / F4; the reference gets it
 * from cuDNN: nn.Conv2d / nn.Conv3d backward under stereoTrainer.py:254-319).
 *   x_hi, x_lo  the fp16 (hi, lo * 2^11) channels-last activation pairs the forward consumed, [N, D, H, W, Cp] halves
 *               (side_ncdhw_to_cl_split_f16; Cp = Cin rounded up to 32, padded channels are zero)
 *   gy          [N, Cout, Do, Ho, Wo] fp32 (NCDHW; 2-D convolutions: D = 1 with the batch as N, or N = 1 with the batch as D)
 *   gw          [Cout][taps][Cq] fp32, Cq = Cp rounded up to 64, OVERWRITTEN (tap-major: permute to [Cout, Cin, taps] and drop
 *               the padded channels)
 *   kernels 1x1x1, 1x3x3, 3x3x3 with padding (k - 1) / 2; stride (H, W) 1 or 2.  Needs Do*Ho*Wo % 8 == 0 and Cout % 8 == 0, else
 *   SIDE_ERR_UNSUPPORTED (the caller keeps cuDNN).  3xFP16 pairs, gy range-scaled by a power of two.
 * --------------------------------------------------------------------------------------------- */
/* Power-of-two range scale of a gradient tensor for the fp16 pairs: s = 2^k with max |x| * s in [2^10, 2^11) (1 for all-zero x);
 * scale_out[0 .. n_scale) = s, inv_out[0 .. n_inv) = 1 / s; scratch_word: 4 bytes of device memory.  No host synchronisation. */
int side_pow2_range_scale(const float *x, long long n, float *scale_out, int n_scale, float *inv_out, int n_inv,
                          void *scratch_word, void *stream);
size_t side_conv_wgrad_tc_ws_bytes(int N, int D, int H, int W, int Cp, int Cout, int kd, int kh, int kw, int stride);
int side_conv_wgrad_tc(const void *x_hi, const void *x_lo, const float *gy, float *gw, int N, int D, int H, int W, int Cp,
                       int Cout, int kd, int kh, int kw, int stride, void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Training-mode BatchNorm (training side of SURVEY.md 8f rows F1 / F4): nn.BatchNorm2d / nn.BatchNorm3d in train() mode
 * (feature_extraction_dla34.py:31-95, stereo_network_old.py:139-171; cuDNN bn_fw_tr / bn_bw kernels in the reference).
 *   x, y, gy, gx [N, C, S] (NCHW / NCDHW, S = product of the spatial sizes); gamma, beta [C] (NULL = 1 / 0)
 *   fwd: batch mean and biased variance per channel (double accumulation), y = gamma (x - mean) / sqrt(var + eps) + beta,
 *        save_mean / save_invstd [C] for the backward, running_mean / running_var (NULL = not tracked) updated with `momentum`
 *        and the unbiased variance, as torch does.
 *   bwd: gx = (gy - sum(gy)/M - (x - mean) invstd^2 sum(gy (x - mean))/M) invstd gamma, ggamma, gbeta (any of the three may be NULL).
 *   N * C <= 65535.  ws: side_bn_train_ws_bytes(N, C, S) bytes, 8-byte aligned.
 * --------------------------------------------------------------------------------------------- */
size_t side_bn_train_ws_bytes(int N, int C, long long S);
int side_bn_train_fwd(const float *x, const float *gamma, const float *beta, float *running_mean, float *running_var,
                      float *y, float *save_mean, float *save_invstd, int N, int C, long long S, float eps, float momentum,
                      void *ws, size_t ws_bytes, void *stream);
int side_bn_train_bwd(const float *x, const float *gy, const float *gamma, const float *save_mean, const float *save_invstd,
                      float *gx, float *ggamma, float *gbeta, int N, int C, long long S, void *ws, size_t ws_bytes,
                      void *stream);

/* Number of kernels launched by this library by the process since the last reset
 * (bench.py's "gpu_launches"). */
long long side_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* SIDE_B200_H_ */
