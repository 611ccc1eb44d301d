// dcn_common.cuh -- sampling geometry shared by the DCNv2 forward / backward kernels.
// Semantics: modulated_deformable_im2col_gpu_kernel + dmcn_im2col_bilinear
// (DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:25-54,125-195), restated in SURVEY.md appendix A.1.
#pragma once
#include "common.cuh"

namespace side {

struct DcnShape {
    int B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dg;
    int Ho, Wo, KK, P;          // KK = kh*kw, P = Ho*Wo
    long long offset_bs, mask_bs;  // batch strides in floats
    int om_cs, om_ps;              // channel / pixel stride of offset and mask (NCHW: P, 1; channels-last rows of ld floats: 1, ld)
    int flags;
};

// One (output pixel, tap) sample: four clamped plane offsets, four bilinear weights (zero for corners that
// fall outside the image or when the whole sample is out of range) and the modulation mask.
struct DcnTap {
    int o1, o2, o3, o4;
    float w1, w2, w3, w4;
    float m;
};

// Also returns the raw fractional parts / validity needed by the backward pass when FULL is set.
struct DcnTapGrad {
    float lh, lw;        // fractional parts
    bool inside;         // h_im > -1 && w_im > -1 && h_im < H && w_im < W
    bool v1, v2, v3, v4; // corner validity
};

__device__ __forceinline__ DcnTap dcn_tap(const DcnShape &s, const float *__restrict__ offset,
                                          const float *__restrict__ mask, int b, int g, int tap, int ho, int wo,
                                          DcnTapGrad *tg = nullptr)
{
    const int p = ho * s.Wo + wo;
    const int i = tap / s.kw, j = tap - i * s.kw;
    const float *ob = offset + (size_t)b * s.offset_bs + ((size_t)g * 2 * s.KK + 2 * tap) * s.om_cs + (size_t)p * s.om_ps;
    const float oh = __ldg(ob), ow = __ldg(ob + s.om_cs);
    float m = __ldg(mask + (size_t)b * s.mask_bs + ((size_t)g * s.KK + tap) * s.om_cs + (size_t)p * s.om_ps);
    if (s.flags & SIDE_DCN_MASK_IS_LOGIT) m = sigmoid_acc(m);
    const float h_im = (float)(ho * s.sh - s.ph + i * s.dh) + oh;
    const float w_im = (float)(wo * s.sw - s.pw + j * s.dw) + ow;
    DcnTap t;
    t.m = m;
    const bool inside = h_im > -1.f && w_im > -1.f && h_im < (float)s.H && w_im < (float)s.W;
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h_low = (int)hf, w_low = (int)wf;
    const int h_high = h_low + 1, w_high = w_low + 1;
    const float lh = h_im - hf, lw = w_im - wf;
    const float hh = 1.f - lh, hw = 1.f - lw;
    const bool v1 = inside && h_low >= 0 && w_low >= 0;
    const bool v2 = inside && h_low >= 0 && w_high <= s.W - 1;
    const bool v3 = inside && h_high <= s.H - 1 && w_low >= 0;
    const bool v4 = inside && h_high <= s.H - 1 && w_high <= s.W - 1;
    const int hl = min(max(h_low, 0), s.H - 1), hhi = min(max(h_high, 0), s.H - 1);
    const int wl = min(max(w_low, 0), s.W - 1), whi = min(max(w_high, 0), s.W - 1);
    t.o1 = hl * s.W + wl;
    t.o2 = hl * s.W + whi;
    t.o3 = hhi * s.W + wl;
    t.o4 = hhi * s.W + whi;
    t.w1 = v1 ? hh * hw : 0.f;
    t.w2 = v2 ? hh * lw : 0.f;
    t.w3 = v3 ? lh * hw : 0.f;
    t.w4 = v4 ? lh * lw : 0.f;
    if (tg) {
        tg->lh = lh; tg->lw = lw; tg->inside = inside;
        tg->v1 = v1; tg->v2 = v2; tg->v3 = v3; tg->v4 = v4;
    }
    return t;
}

struct DcnFwdArgs {
    const float *x, *offset, *mask, *wt, *bias, *scale, *shift;
    float *y;
    DcnShape s;
};

inline int dcn_fill_shape(DcnShape &s, int B, int Cin, int H, int W, int Cout, int kh, int kw, int sh, int sw, int ph,
                          int pw, int dh, int dw, int dg, long long offset_bs, long long mask_bs, int flags)
{
    SIDE_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, "dcn: bad tensor shape");
    SIDE_REQUIRE(kh > 0 && kw > 0 && sh > 0 && sw > 0 && ph >= 0 && pw >= 0 && dh > 0 && dw > 0, "dcn: bad conv params");
    SIDE_REQUIRE(dg >= 1 && Cin % dg == 0, "dcn: deformable_groups=%d must divide Cin=%d", dg, Cin);
    s.B = B; s.Cin = Cin; s.H = H; s.W = W; s.Cout = Cout; s.kh = kh; s.kw = kw; s.sh = sh; s.sw = sw;
    s.ph = ph; s.pw = pw; s.dh = dh; s.dw = dw; s.dg = dg; s.flags = flags;
    s.Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
    s.Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
    SIDE_REQUIRE(s.Ho > 0 && s.Wo > 0, "dcn: empty output");
    s.KK = kh * kw;
    s.P = s.Ho * s.Wo;
    s.offset_bs = offset_bs ? offset_bs : (long long)dg * 2 * s.KK * s.P;
    s.mask_bs = mask_bs ? mask_bs : (long long)dg * s.KK * s.P;
    s.om_cs = s.P; s.om_ps = 1;
    SIDE_REQUIRE((long long)B * s.P < (1ll << 31) && (long long)Cin * H * W < (1ll << 31), "dcn: tensor too large");
    return SIDE_OK;
}

}  // namespace side
