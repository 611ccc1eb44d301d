/*
 * side_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into / called by the product).
 *
 * Plain-C, single-thread CPU restatement of the arithmetic on SIDE's stereo hot path
 * (SURVEY.md section 8a).  Each function cites the reference file:line it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load the shared object built from this file.
 *
 * Pinning (see DESIGN.md "Oracle"):
 *   - orc_dcn_*           : pinned against the reference's own DCNv2 KAT (DCNv2/test.py:32-67,
 *                           zero-offset identity) and against the reference Python wrapper
 *                           (DCNv2/dcn_v2.py) executed in the build container with _ext mapped to
 *                           torchvision.ops.deform_conv2d (tests/golden/dcn npz files).
 *   - orc_roi_align       : third-party arithmetic -- torchvision.ops.roi_align (not under
 *                           /root/reference; version unpinned by the reference, container has
 *                           torchvision 0.26.0).  Restates the published legacy (aligned=False)
 *                           algorithm; pinned against torchvision 0.26.0 CPU outputs and against the
 *                           reference call sites stereo_network_old.py:271,372-376 (golden vectors).
 *   - orc_proposal_shift, orc_xcross_gate, orc_softargmin, orc_nms_topk, orc_*_decode :
 *                           pinned against the reference's own Python (stereo_network_old.py,
 *                           decode.py) executed on seeded inputs (tests/golden/ npz files).
 *   - orc_concat_volume / orc_gwc_volume : NO reference call site exists --
 *                           PARITY UNPINNED for these two; semantics fixed in SURVEY.md section 8 row A6.
 *
 * Compile: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).  -ffp-contract=off
 * matters: the reference's arithmetic is a chain of separately rounded float32 torch ops.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * A2: DCNv2 forward.
 * dmcn_im2col_bilinear: DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:25-54
 * ------------------------------------------------------------------------------------------ */
static float dcn_bilinear(const float *im, int data_width, int height, int width, float h, float w)
{
    int h_low = (int)floorf(h);
    int w_low = (int)floorf(w);
    int h_high = h_low + 1;
    int w_high = w_low + 1;
    float lh = h - h_low;
    float lw = w - w_low;
    float hh = 1 - lh, hw = 1 - lw;
    float v1 = 0, v2 = 0, v3 = 0, v4 = 0;
    if (h_low >= 0 && w_low >= 0) v1 = im[h_low * data_width + w_low];
    if (h_low >= 0 && w_high <= width - 1) v2 = im[h_low * data_width + w_high];
    if (h_high <= height - 1 && w_low >= 0) v3 = im[h_high * data_width + w_low];
    if (h_high <= height - 1 && w_high <= width - 1) v4 = im[h_high * data_width + w_high];
    float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    return (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
}

/* modulated_deformable_im2col_gpu_kernel: dcn_v2_im2col_cuda.cu:125-195 for ONE sample.
 * col layout [Cin*kh*kw, Ho*Wo] (the per-sample slice of the reference's `columns`). */
static void dcn_im2col_sample(const float *im, const float *off, const float *msk, int Cin, int H, int W,
                              int kh, int kw, int ph, int pw, int sh, int sw, int dh, int dw, int dg,
                              int Ho, int Wo, float *col)
{
    const int cpg = Cin / dg;
    for (int c = 0; c < Cin; ++c) {
        const int g = c / cpg;
        const float *imc = im + (size_t)c * H * W;
        const float *offg = off + (size_t)g * 2 * kh * kw * Ho * Wo;
        const float *mskg = msk + (size_t)g * kh * kw * Ho * Wo;
        for (int ho = 0; ho < Ho; ++ho)
            for (int wo = 0; wo < Wo; ++wo) {
                const int h_in = ho * sh - ph, w_in = wo * sw - pw;
                for (int i = 0; i < kh; ++i)
                    for (int j = 0; j < kw; ++j) {
                        const int t = i * kw + j;
                        const float oh = offg[((size_t)(2 * t) * Ho + ho) * Wo + wo];
                        const float ow = offg[((size_t)(2 * t + 1) * Ho + ho) * Wo + wo];
                        const float m = mskg[((size_t)t * Ho + ho) * Wo + wo];
                        float val = 0.f;
                        const float h_im = h_in + i * dh + oh;
                        const float w_im = w_in + j * dw + ow;
                        if (h_im > -1 && w_im > -1 && h_im < H && w_im < W)
                            val = dcn_bilinear(imc, W, H, W, h_im, w_im);
                        col[((size_t)(c * kh * kw + t) * Ho + ho) * Wo + wo] = val * m;
                    }
            }
    }
}

/* dcn_v2_cuda_forward: DCNv2/src/cuda/dcn_v2_cuda.cu:43-173.
 * out[b] = ones*bias (rank-1 GEMM :124-138) then += W[Cout, Cin*kh*kw] * columns[b] (:150-164).
 * The cuBLAS summation order is unspecified; this restatement accumulates in double, in k order. */
ORC_API int orc_dcn_forward(const float *x, const float *offset, const float *mask, const float *w,
                            const float *bias, float *y, int B, int Cin, int H, int W, int Cout, int kh,
                            int kw, int sh, int sw, int ph, int pw, int dh, int dw, int dg)
{
    const int Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
    const int Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
    const int K = Cin * kh * kw, P = Ho * Wo;
    float *col = (float *)malloc(sizeof(float) * (size_t)K * P);
    double *acc = (double *)malloc(sizeof(double) * (size_t)P);
    if (!col || !acc) return -1;
    for (int b = 0; b < B; ++b) {
        dcn_im2col_sample(x + (size_t)b * Cin * H * W, offset + (size_t)b * dg * 2 * kh * kw * P,
                          mask + (size_t)b * dg * kh * kw * P, Cin, H, W, kh, kw, ph, pw, sh, sw, dh, dw, dg, Ho,
                          Wo, col);
        for (int o = 0; o < Cout; ++o) {
            for (int p = 0; p < P; ++p) acc[p] = bias ? bias[o] : 0.0;
            for (int k = 0; k < K; ++k) {
                const double wk = w[(size_t)o * K + k];
                const float *ck = col + (size_t)k * P;
                for (int p = 0; p < P; ++p) acc[p] += wk * ck[p];
            }
            float *yo = y + ((size_t)b * Cout + o) * P;
            for (int p = 0; p < P; ++p) yo[p] = (float)acc[p];
        }
    }
    free(col);
    free(acc);
    return 0;
}

/* dmcn_get_gradient_weight: dcn_v2_im2col_cuda.cu:56-80 */
static float dcn_grad_weight(float ah, float aw, int h, int w, int height, int width)
{
    if (ah <= -1 || ah >= height || aw <= -1 || aw >= width) return 0;
    int hl = (int)floorf(ah), wl = (int)floorf(aw);
    int hh = hl + 1, wh = wl + 1;
    float weight = 0;
    if (h == hl && w == wl) weight = (h + 1 - ah) * (w + 1 - aw);
    if (h == hl && w == wh) weight = (h + 1 - ah) * (aw + 1 - w);
    if (h == hh && w == wl) weight = (ah + 1 - h) * (w + 1 - aw);
    if (h == hh && w == wh) weight = (ah + 1 - h) * (aw + 1 - w);
    return weight;
}

/* dmcn_get_coordinate_weight: dcn_v2_im2col_cuda.cu:82-123 */
static float dcn_coord_weight(float ah, float aw, int height, int width, const float *im, int data_width,
                              int bp_dir)
{
    if (ah <= -1 || ah >= height || aw <= -1 || aw >= width) return 0;
    int hl = (int)floorf(ah), wl = (int)floorf(aw);
    int hh = hl + 1, wh = wl + 1;
    float weight = 0;
    if (bp_dir == 0) {
        if (hl >= 0 && wl >= 0) weight += -1 * (wl + 1 - aw) * im[hl * data_width + wl];
        if (hl >= 0 && wh <= width - 1) weight += -1 * (aw - wl) * im[hl * data_width + wh];
        if (hh <= height - 1 && wl >= 0) weight += (wl + 1 - aw) * im[hh * data_width + wl];
        if (hh <= height - 1 && wh <= width - 1) weight += (aw - wl) * im[hh * data_width + wh];
    } else {
        if (hl >= 0 && wl >= 0) weight += -1 * (hl + 1 - ah) * im[hl * data_width + wl];
        if (hl >= 0 && wh <= width - 1) weight += (hl + 1 - ah) * im[hl * data_width + wh];
        if (hh <= height - 1 && wl >= 0) weight += -1 * (ah - hl) * im[hh * data_width + wl];
        if (hh <= height - 1 && wh <= width - 1) weight += (ah - hl) * im[hh * data_width + wh];
    }
    return weight;
}

/* dcn_v2_cuda_backward: dcn_v2_cuda.cu:207-336 (serial loop over the batch :260), with
 * modulated_deformable_col2im_coord_gpu_kernel (dcn_v2_im2col_cuda.cu:256-327),
 * modulated_deformable_col2im_gpu_kernel (:197-254), im2col recompute + the two Sgemm and the Sgemv.
 * Sums are accumulated in double (the reference's atomicAdd order is unspecified anyway). */
ORC_API int orc_dcn_backward(const float *x, const float *offset, const float *mask, const float *w,
                             const float *gy, float *gx, float *goff, float *gmask, float *gw, float *gb,
                             int B, int Cin, int H, int W, int Cout, int kh, int kw, int sh, int sw, int ph,
                             int pw, int dh, int dw, int dg)
{
    const int Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
    const int Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
    const int KK = kh * kw, K = Cin * KK, P = Ho * Wo, cpg = Cin / dg;
    float *col = (float *)malloc(sizeof(float) * (size_t)K * P);
    double *gcol = (double *)malloc(sizeof(double) * (size_t)K * P);
    double *gxd = (double *)calloc((size_t)B * Cin * H * W, sizeof(double));
    double *gwd = (double *)calloc((size_t)Cout * K, sizeof(double));
    double *gbd = (double *)calloc((size_t)Cout, sizeof(double));
    if (!col || !gcol || !gxd || !gwd || !gbd) return -1;
    for (int b = 0; b < B; ++b) {
        const float *xb = x + (size_t)b * Cin * H * W;
        const float *offb = offset + (size_t)b * dg * 2 * KK * P;
        const float *mskb = mask + (size_t)b * dg * KK * P;
        const float *gyb = gy + (size_t)b * Cout * P;
        /* columns = W^T * grad_output  (:274-277) */
        for (int k = 0; k < K; ++k) {
            double *gk = gcol + (size_t)k * P;
            for (int p = 0; p < P; ++p) gk[p] = 0.0;
            for (int o = 0; o < Cout; ++o) {
                const double wk = w[(size_t)o * K + k];
                const float *g = gyb + (size_t)o * P;
                for (int p = 0; p < P; ++p) gk[p] += wk * g[p];
            }
        }
        /* grad offset / mask (col2im_coord) */
        for (int g = 0; g < dg; ++g)
            for (int t = 0; t < KK; ++t) {
                const int i = t / kw, j = t % kw;
                for (int ho = 0; ho < Ho; ++ho)
                    for (int wo = 0; wo < Wo; ++wo) {
                        const int p = ho * Wo + wo;
                        const float oh = offb[((size_t)(g * 2 * KK + 2 * t)) * P + p];
                        const float ow = offb[((size_t)(g * 2 * KK + 2 * t + 1)) * P + p];
                        const float m = mskb[((size_t)(g * KK + t)) * P + p];
                        float inv_h = ho * sh - ph + i * dh + oh;
                        float inv_w = wo * sw - pw + j * dw + ow;
                        int inside = !(inv_h <= -1 || inv_w <= -1 || inv_h >= H || inv_w >= W);
                        double vh = 0, vw = 0, mv = 0;
                        for (int cc = 0; cc < cpg; ++cc) {
                            const int c = g * cpg + cc;
                            const float *imc = xb + (size_t)c * H * W;
                            const double gc = gcol[(size_t)(c * KK + t) * P + p];
                            if (inside) {
                                mv += gc * dcn_bilinear(imc, W, H, W, inv_h, inv_w);
                                vh += (double)dcn_coord_weight(inv_h, inv_w, H, W, imc, W, 0) * gc * m;
                                vw += (double)dcn_coord_weight(inv_h, inv_w, H, W, imc, W, 1) * gc * m;
                            }
                        }
                        goff[((size_t)b * dg * 2 * KK + g * 2 * KK + 2 * t) * P + p] = (float)vh;
                        goff[((size_t)b * dg * 2 * KK + g * 2 * KK + 2 * t + 1) * P + p] = (float)vw;
                        gmask[((size_t)b * dg * KK + g * KK + t) * P + p] = (float)mv;
                    }
            }
        /* grad input (col2im): scatter over the 5x5 window exactly like :238-252 */
        for (int c = 0; c < Cin; ++c) {
            const int g = c / cpg;
            for (int t = 0; t < KK; ++t) {
                const int i = t / kw, j = t % kw;
                for (int ho = 0; ho < Ho; ++ho)
                    for (int wo = 0; wo < Wo; ++wo) {
                        const int p = ho * Wo + wo;
                        const float oh = offb[((size_t)(g * 2 * KK + 2 * t)) * P + p];
                        const float ow = offb[((size_t)(g * 2 * KK + 2 * t + 1)) * P + p];
                        const float m = mskb[((size_t)(g * KK + t)) * P + p];
                        const float ch = ho * sh - ph + i * dh + oh;
                        const float cw = wo * sw - pw + j * dw + ow;
                        const double top = gcol[(size_t)(c * KK + t) * P + p] * m;
                        const int cur_h = (int)ch, cur_w = (int)cw;
                        for (int dy = -2; dy <= 2; ++dy)
                            for (int dx = -2; dx <= 2; ++dx)
                                if (cur_h + dy >= 0 && cur_h + dy < H && cur_w + dx >= 0 && cur_w + dx < W &&
                                    fabsf(ch - (cur_h + dy)) < 1 && fabsf(cw - (cur_w + dx)) < 1) {
                                    float wt = dcn_grad_weight(ch, cw, cur_h + dy, cur_w + dx, H, W);
                                    gxd[(((size_t)b * Cin + c) * H + cur_h + dy) * W + cur_w + dx] += wt * top;
                                }
                    }
            }
        }
        /* grad weight (:303-320) and grad bias (:325-330) */
        dcn_im2col_sample(xb, offb, mskb, Cin, H, W, kh, kw, ph, pw, sh, sw, dh, dw, dg, Ho, Wo, col);
        for (int o = 0; o < Cout; ++o) {
            const float *g = gyb + (size_t)o * P;
            for (int k = 0; k < K; ++k) {
                const float *ck = col + (size_t)k * P;
                double s = 0;
                for (int p = 0; p < P; ++p) s += (double)g[p] * ck[p];
                gwd[(size_t)o * K + k] += s;
            }
            double s = 0;
            for (int p = 0; p < P; ++p) s += g[p];
            gbd[o] += s;
        }
    }
    for (size_t i = 0; i < (size_t)B * Cin * H * W; ++i) gx[i] = (float)gxd[i];
    for (size_t i = 0; i < (size_t)Cout * K; ++i) gw[i] = (float)gwd[i];
    for (int o = 0; o < Cout; ++o) gb[o] = (float)gbd[o];
    free(col); free(gcol); free(gxd); free(gwd); free(gbd);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A5: RoIAlign, legacy (aligned=False), as torchvision.ops.roi_align computes it on CPU
 * (third-party; call site stereo_network_old.py:271 RoIAlign((P,P), spatial_scale=1, sampling_ratio=2)).
 * One sample point; returns the 4 separately rounded float products summed left to right.
 * ------------------------------------------------------------------------------------------ */
static float roi_bilinear(const float *im, int H, int W, float y, float x)
{
    if (y < -1.0f || y > H || x < -1.0f || x > W) return 0.f;
    if (y <= 0) y = 0;
    if (x <= 0) x = 0;
    int y_low = (int)y, x_low = (int)x, y_high, x_high;
    if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
    if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
    float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
    float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
    return w1 * im[y_low * W + x_low] + w2 * im[y_low * W + x_high] + w3 * im[y_high * W + x_low] +
           w4 * im[y_high * W + x_high];
}

/* out[C, P, P] (channel stride out_cs floats) for one RoI (b, x1, y1, x2, y2) */
static void roi_align_one(const float *feat, int C, int H, int W, const float *roi, float spatial_scale,
                          int sampling_ratio, int P, float *out, size_t out_cs)
{
    const int b = (int)roi[0];
    const float x1 = roi[1] * spatial_scale, y1 = roi[2] * spatial_scale;
    const float x2 = roi[3] * spatial_scale, y2 = roi[4] * spatial_scale;
    float roi_w = x2 - x1, roi_h = y2 - y1;
    roi_w = roi_w > 1.f ? roi_w : 1.f;
    roi_h = roi_h > 1.f ? roi_h : 1.f;
    const float bin_h = roi_h / (float)P, bin_w = roi_w / (float)P;
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_h / P);
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_w / P);
    const float count = (float)(gh * gw > 1 ? gh * gw : 1);
    for (int c = 0; c < C; ++c) {
        const float *im = feat + ((size_t)b * C + c) * H * W;
        for (int ph = 0; ph < P; ++ph)
            for (int pw = 0; pw < P; ++pw) {
                float acc = 0.f;
                for (int iy = 0; iy < gh; ++iy) {
                    const float yy = y1 + ph * bin_h + (float)(iy + .5f) * bin_h / (float)gh;
                    for (int ix = 0; ix < gw; ++ix) {
                        const float xx = x1 + pw * bin_w + (float)(ix + .5f) * bin_w / (float)gw;
                        acc += roi_bilinear(im, H, W, yy, xx);
                    }
                }
                out[(size_t)c * out_cs + ph * P + pw] = acc / count;
            }
    }
}

ORC_API int orc_roi_align(const float *feat, const float *rois, float *out, int N, int C, int H, int W, int P,
                          float spatial_scale, int sampling_ratio)
{
    for (int n = 0; n < N; ++n)
        roi_align_one(feat, C, H, W, rois + (size_t)n * 5, spatial_scale, sampling_ratio, P,
                      out + (size_t)n * C * P * P, (size_t)P * P);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A4: get_proposal_shift -- stereo_network_old.py:34-133.
 * Every line below is ONE float32 torch op in the reference (no contraction possible there).
 * Boxes must already be grouped by image in ascending b (the reference's output order, :45-82,:125-131);
 * returns the number of RoIs.  x_clamp = input_w//4 - 1 = 319 (:21,:70).
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_proposal_shift(const float *left, const float *right, const float *fbs, int N, int D,
                               float x_clamp, float *pro_left, float *pro_right, float *depth_bin)
{
    const float depth_max = 87.0f;
    for (int n = 0; n < N; ++n) {
        const float *l = left + (size_t)n * 5, *r = right + (size_t)n * 5;
        const int b = (int)l[0];
        const float fb = fbs[b];
        const float xmin = fminf(l[1], r[1]), ymin = fminf(l[2], r[2]);
        const float xmax = fmaxf(l[3], r[3]), ymax = fmaxf(l[4], r[4]);
        float t = xmax - xmin;       /* :58 (xmax - xmin)            */
        t = t * 0.9f;                /*     *0.9                     */
        t = t * 4.0f;                /*     *4                       */
        float dmin = fb / t;         /*     fb/(...)                 */
        dmin = fminf(fmaxf(dmin, 1.0f), 87.0f); /* :59 clamp (NaN cannot occur: fb>0) */
        for (int i = 0; i < D; ++i) {
            const float rate = (float)((double)i / (double)(D - 1)); /* :38-39 python double -> float32 */
            float u = depth_max - dmin;                               /* :60 */
            u = u * rate;
            const float dbin = depth_max - u;
            float disp = fb / dbin;                                   /* :61 */
            disp = disp / 8.0f;
            depth_bin[(size_t)n * D + i] = dbin;
            float *pl = pro_left + ((size_t)i * N + n) * 5, *pr = pro_right + ((size_t)i * N + n) * 5;
            pl[0] = l[0]; pl[1] = fminf(xmin + disp, x_clamp); pl[2] = ymin; /* :70-72 */
            pl[3] = fminf(xmax + disp, x_clamp); pl[4] = ymax;
            pr[0] = l[0]; pr[1] = fmaxf(xmin - disp, 0.f); pr[2] = ymin;     /* :75-77 */
            pr[3] = fmaxf(xmax - disp, 0.f); pr[4] = ymax;
        }
    }
    return N;
}

/* ------------------------------------------------------------------------------------------
 * A5: instance cost volume -- stereo_network_old.py:368-376.
 * cost[n, 0:C, i] = RoIAlign(featL, pro_left[i,n]); cost[n, C:2C, i] = RoIAlign(featR, pro_right[i,n]);
 * cost[n, 2C:3C, i] = L - R.     cost layout [N, 3C, D, P, P].
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_inst_costvol(const float *featL, const float *featR, const float *pro_left,
                             const float *pro_right, float *cost, int N, int C, int H, int W, int D, int P,
                             int sampling_ratio)
{
    const size_t cs = (size_t)D * P * P; /* channel stride inside one RoI's volume */
    for (int n = 0; n < N; ++n)
        for (int i = 0; i < D; ++i) {
            float *base = cost + (size_t)n * 3 * C * cs + (size_t)i * P * P;
            roi_align_one(featL, C, H, W, pro_left + ((size_t)i * N + n) * 5, 1.0f, sampling_ratio, P, base, cs);
            roi_align_one(featR, C, H, W, pro_right + ((size_t)i * N + n) * 5, 1.0f, sampling_ratio, P,
                          base + (size_t)C * cs, cs);
            for (int c = 0; c < C; ++c)
                for (int q = 0; q < P * P; ++q)
                    base[(size_t)(2 * C + c) * cs + q] = base[(size_t)c * cs + q] - base[(size_t)(C + c) * cs + q];
        }
    return 0;
}

/* A6: cosine gate x_cross -- stereo_network_old.py:197-203 (num_channels generalised from 32 to C).
 * xc[n,i] = sum(L*R) / max(sqrt(sum L^2) * sqrt(sum R^2), 0.01);  cost[n,:,i] *= xc[n,i].
 * torch.sum's float32 reduction order is unspecified -> double accumulators here (tolerance 1e-4). */
ORC_API int orc_xcross_gate(const float *cost, float *out, float *xc, int N, int C, int D, int P)
{
    const size_t cs = (size_t)D * P * P, PP = (size_t)P * P;
    for (int n = 0; n < N; ++n)
        for (int i = 0; i < D; ++i) {
            const float *base = cost + (size_t)n * 3 * C * cs + (size_t)i * PP;
            double sl = 0, sr = 0, slr = 0;
            for (int c = 0; c < C; ++c)
                for (size_t q = 0; q < PP; ++q) {
                    const double l = base[(size_t)c * cs + q], r = base[(size_t)(C + c) * cs + q];
                    sl += l * l; sr += r * r; slr += l * r;
                }
            float den = sqrtf((float)sl) * sqrtf((float)sr);
            den = den > 0.01f ? den : 0.01f;
            const float g = (float)slr / den;
            if (xc) xc[(size_t)n * D + i] = g;
            if (out) {
                float *ob = out + (size_t)n * 3 * C * cs + (size_t)i * PP;
                for (int c = 0; c < 3 * C; ++c)
                    for (size_t q = 0; q < PP; ++q) ob[(size_t)c * cs + q] = base[(size_t)c * cs + q] * g;
            }
        }
    return 0;
}

/* A7 tail: AvgPool2d(S,S) -> softmax over D -> sum_i p_i * depth_bin_i -- stereo_network_old.py:228-236.
 * logits layout [N, D, S, S] (the squeezed classify output, S=4). */
ORC_API int orc_softargmin(const float *logits, const float *depth_bin, float *depth, float *prob, int N,
                           int D, int S)
{
    float *lg = (float *)malloc(sizeof(float) * (size_t)D);
    if (!lg) return -1;
    for (int n = 0; n < N; ++n) {
        float mx = -INFINITY;
        for (int i = 0; i < D; ++i) {
            float s = 0.f;
            for (int q = 0; q < S * S; ++q) s += logits[((size_t)n * D + i) * S * S + q];
            lg[i] = s / (float)(S * S);
            mx = lg[i] > mx ? lg[i] : mx;
        }
        double den = 0;
        for (int i = 0; i < D; ++i) den += exp((double)lg[i] - mx);
        float acc = 0.f;
        for (int i = 0; i < D; ++i) {
            const float p = (float)(exp((double)lg[i] - mx) / den);
            if (prob) prob[(size_t)n * D + i] = p;
            acc += p * depth_bin[(size_t)n * D + i]; /* :234-235 sequential float32 accumulation */
        }
        depth[n] = acc;
    }
    free(lg);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A8: _nms + _topk -- decode.py:9-33.  Tie rule (torch.topk leaves it unspecified):
 * larger value first, then LOWER flat index first.  heat_is_logit=1 applies sigmoid first (bbox_decode :93).
 * Outputs per image: score[K], ind[K] (flat y*W+x), cls[K], ys[K], xs[K].
 * ------------------------------------------------------------------------------------------ */
typedef struct { float v; int i; } orc_pair;
static int pair_cmp(const void *a, const void *b)
{
    const orc_pair *p = (const orc_pair *)a, *q = (const orc_pair *)b;
    if (p->v > q->v) return -1;
    if (p->v < q->v) return 1;
    return p->i < q->i ? -1 : (p->i > q->i ? 1 : 0);
}
static float orc_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

ORC_API int orc_nms_topk(const float *heat, int B, int Cat, int H, int W, int K, int heat_is_logit,
                         float *score, int *ind, int *cls, float *ys, float *xs)
{
    const int HW = H * W;
    if (K > HW) return -2;
    float *hs = (float *)malloc(sizeof(float) * (size_t)HW);
    orc_pair *pp = (orc_pair *)malloc(sizeof(orc_pair) * (size_t)HW);
    orc_pair *cand = (orc_pair *)malloc(sizeof(orc_pair) * (size_t)Cat * K);
    int *cand_ind = (int *)malloc(sizeof(int) * (size_t)Cat * K);
    if (!hs || !pp || !cand || !cand_ind) return -1;
    for (int b = 0; b < B; ++b) {
        for (int c = 0; c < Cat; ++c) {
            const float *h = heat + ((size_t)b * Cat + c) * HW;
            for (int q = 0; q < HW; ++q) hs[q] = heat_is_logit ? orc_sigmoid(h[q]) : h[q];
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    float m = -INFINITY; /* max_pool2d pads with -inf (:12-13) */
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int yy = y + dy, xx = x + dx;
                            if (yy >= 0 && yy < H && xx >= 0 && xx < W && hs[yy * W + xx] > m) m = hs[yy * W + xx];
                        }
                    const float keep = (m == hs[y * W + x]) ? 1.f : 0.f; /* :14 */
                    pp[y * W + x].v = hs[y * W + x] * keep;              /* :15 */
                    pp[y * W + x].i = y * W + x;
                }
            qsort(pp, (size_t)HW, sizeof(orc_pair), pair_cmp);
            for (int k = 0; k < K; ++k) {
                cand[c * K + k].v = pp[k].v;
                cand[c * K + k].i = c * K + k; /* position in the [Cat*K] view (:26) */
                cand_ind[c * K + k] = pp[k].i;
            }
        }
        qsort(cand, (size_t)Cat * K, sizeof(orc_pair), pair_cmp);
        for (int k = 0; k < K; ++k) {
            const int pos = cand[k].i;
            const int id = cand_ind[pos];
            score[(size_t)b * K + k] = cand[k].v;
            ind[(size_t)b * K + k] = id;
            cls[(size_t)b * K + k] = pos / K;            /* :27 */
            ys[(size_t)b * K + k] = (float)(id / W);     /* :23 */
            xs[(size_t)b * K + k] = (float)(id % W);     /* :24 */
        }
    }
    free(hs); free(pp); free(cand); free(cand_ind);
    return 0;
}

/* bbox_decode -- decode.py:91-126.  Writes the UNCOMPACTED [B,K,5] boxes plus keep[B*K] (:123);
 * the caller compacts (boolean indexing).  wh_scale multiplies wh first (stereo_network_old.py:360). */
ORC_API int orc_bbox_decode(const float *hm, const float *wh, const float *reg, int B, int Cat, int H, int W,
                            int K, float wh_scale, float *bbox, float *bbox_right, uint8_t *keep)
{
    const int HW = H * W;
    float *score = (float *)malloc(sizeof(float) * (size_t)B * K * 3);
    int *ind = (int *)malloc(sizeof(int) * (size_t)B * K * 2);
    if (!score || !ind) return -1;
    float *ys = score + (size_t)B * K, *xs = ys + (size_t)B * K;
    int *cls = ind + (size_t)B * K;
    int rc = orc_nms_topk(hm, B, Cat, H, W, K, 1, score, ind, cls, ys, xs);
    if (rc) return rc;
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < K; ++k) {
            const int id = ind[(size_t)b * K + k];
            float r[3], w[3];
            for (int j = 0; j < 3; ++j) {
                r[j] = reg[((size_t)b * 3 + j) * HW + id];
                w[j] = wh[((size_t)b * 3 + j) * HW + id] * wh_scale;
            }
            const float x = xs[(size_t)b * K + k], y = ys[(size_t)b * K + k];
            const float cx = x + r[0], cxr = x + r[1], cy = y + r[2]; /* :101-103 */
            float *o = bbox + ((size_t)b * K + k) * 5, *q = bbox_right + ((size_t)b * K + k) * 5;
            o[0] = (float)b; o[1] = cx - 0.5f * w[0]; o[2] = cy - 0.5f * w[2]; /* :113-114 */
            o[3] = cx + 0.5f * w[0]; o[4] = cy + 0.5f * w[2];
            q[0] = (float)b; q[1] = cxr - 0.5f * w[1]; q[2] = cy - 0.5f * w[2]; /* :116-117 */
            q[3] = cxr + 0.5f * w[1]; q[4] = cy + 0.5f * w[2];
            const float s = ((o[1] + o[2]) + o[3]) + o[4]; /* torch.sum over 4 floats (:122) */
            keep[(size_t)b * K + k] = s > 0 ? 1 : 0;
        }
    free(score); free(ind);
    return 0;
}

static int argmax_first(const float *v, size_t stride, int n)
{
    int best = 0;
    float bv = v[0];
    for (int i = 1; i < n; ++i)
        if (v[(size_t)i * stride] > bv) { bv = v[(size_t)i * stride]; best = i; }
    return best;
}

/* ddd_decode -- decode.py:35-89.  heat is already sigmoid-ed by the caller (stereoDetector.py:88).
 * detections [B,K,6], detections_right [B,K,6], info_3d [B,K,9].
 * kept_type = floor(argmax/grid) (the reference relied on torch-0.4 integer division). */
ORC_API int orc_ddd_decode(const float *heat, const float *kept, const float *dim, const float *orien,
                           const float *wh, const float *reg, int B, int Cat, int H, int W, int grid, int K,
                           float *det, float *det_right, float *info)
{
    const int HW = H * W;
    float *score = (float *)malloc(sizeof(float) * (size_t)B * K * 3);
    int *ind = (int *)malloc(sizeof(int) * (size_t)B * K * 2);
    if (!score || !ind) return -1;
    float *ys = score + (size_t)B * K, *xs = ys + (size_t)B * K;
    int *cls = ind + (size_t)B * K;
    int rc = orc_nms_topk(heat, B, Cat, H, W, K, 0, score, ind, cls, ys, xs);
    if (rc) return rc;
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < K; ++k) {
            const size_t bk = (size_t)b * K + k;
            const int id = ind[bk];
            float r[3], w[3];
            for (int j = 0; j < 3; ++j) {
                r[j] = reg[((size_t)b * 3 + j) * HW + id];
                w[j] = wh[((size_t)b * 3 + j) * HW + id];
            }
            float *d = det + bk * 6, *dr = det_right + bk * 6, *f = info + bk * 9;
            d[0] = xs[bk] + r[0]; d[1] = ys[bk] + r[2]; d[2] = w[0]; d[3] = w[2]; d[4] = score[bk]; d[5] = (float)cls[bk];
            dr[0] = xs[bk] + r[1]; dr[1] = ys[bk] + r[2]; dr[2] = w[1]; dr[3] = w[2]; dr[4] = score[bk]; dr[5] = (float)cls[bk];
            for (int j = 0; j < 3; ++j) f[j] = dim[((size_t)b * 3 + j) * HW + id];
            for (int j = 0; j < 2; ++j) f[3 + j] = orien[((size_t)b * 2 + j) * HW + id];
            const float *kp = kept + (size_t)b * 6 * grid * HW + id;
            const int a0 = argmax_first(kp, (size_t)HW, 4 * grid);                      /* :60-63 */
            const int a1 = argmax_first(kp + (size_t)4 * grid * HW, (size_t)HW, grid);  /* :67-70 */
            const int a2 = argmax_first(kp + (size_t)5 * grid * HW, (size_t)HW, grid);  /* :72-75 */
            f[5] = (float)a1; f[6] = (float)a2; f[7] = (float)(a0 % grid); f[8] = (float)(a0 / grid); /* :81-82 */
        }
    free(score); free(ind);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A6 generic full-image builders (PSMNet / GwcNet convention; PARITY UNPINNED -- no reference call site).
 * concat[b, 0:C, d, y, x] = L[b,:,y,x]*[x>=d];  concat[b, C:2C, d, y, x] = R[b,:,y,x-d]*[x>=d]
 * gwc[b, g, d, y, x] = (1/(C/G)) * sum_{c in g} L[b,c,y,x]*R[b,c,y,x-d]*[x>=d]
 * ------------------------------------------------------------------------------------------ */
ORC_API int orc_concat_volume(const float *L, const float *R, float *vol, int B, int C, int H, int W, int D)
{
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int d = 0; d < D; ++d)
                for (int y = 0; y < H; ++y) {
                    const float *l = L + (((size_t)b * C + c) * H + y) * W, *r = R + (((size_t)b * C + c) * H + y) * W;
                    float *vl = vol + ((((size_t)b * 2 * C + c) * D + d) * H + y) * W;
                    float *vr = vol + ((((size_t)b * 2 * C + C + c) * D + d) * H + y) * W;
                    for (int x = 0; x < W; ++x) {
                        vl[x] = x >= d ? l[x] : 0.f;
                        vr[x] = x >= d ? r[x - d] : 0.f;
                    }
                }
    return 0;
}

ORC_API int orc_gwc_volume(const float *L, const float *R, float *vol, int B, int C, int H, int W, int D, int G)
{
    const int cpg = C / G;
    for (int b = 0; b < B; ++b)
        for (int g = 0; g < G; ++g)
            for (int d = 0; d < D; ++d)
                for (int y = 0; y < H; ++y) {
                    float *v = vol + ((((size_t)b * G + g) * D + d) * H + y) * W;
                    for (int x = 0; x < W; ++x) {
                        double s = 0;
                        if (x >= d)
                            for (int c = g * cpg; c < (g + 1) * cpg; ++c)
                                s += (double)L[(((size_t)b * C + c) * H + y) * W + x] *
                                     R[(((size_t)b * C + c) * H + y) * W + x - d];
                        v[x] = (float)(s / cpg);
                    }
                }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Aggregation network layer: the arithmetic behind
 * cost_volume.dres0/dres1/dres2/classify (stereo_network_old.py:139-171, convbn_3d :29-32):
 *   y = relu?( conv3d(x, w; 3x3x3, stride 1, padding 1, no bias) * scale[o] + shift[o] ) (+ residual)
 * where scale/shift are the eval-mode BatchNorm3d folded (gamma / sqrt(var + eps), beta - mean * scale).
 * nn.Conv3d itself is a third-party (ATen / cuDNN) op: this restatement is pinned against torch's CPU conv3d in
 * tests/test_oracle_golden.py.  NCDHW layout, double accumulation.
 * --------------------------------------------------------------------------------------------- */
ORC_API int orc_conv3d_bn_relu(const float *x, const float *w, const float *scale, const float *shift,
                               const float *residual, float *y, int N, int Cin, int D, int H, int W, int Cout, int relu)
{
    const size_t S = (size_t)D * H * W;
    for (int n = 0; n < N; ++n)
        for (int o = 0; o < Cout; ++o)
            for (int d = 0; d < D; ++d)
                for (int h = 0; h < H; ++h)
                    for (int xw = 0; xw < W; ++xw) {
                        double acc = 0;
                        for (int c = 0; c < Cin; ++c) {
                            const float *xp = x + ((size_t)n * Cin + c) * S;
                            const float *wp = w + ((size_t)o * Cin + c) * 27;
                            for (int kd = 0; kd < 3; ++kd) {
                                const int dd = d + kd - 1;
                                if (dd < 0 || dd >= D) continue;
                                for (int kh = 0; kh < 3; ++kh) {
                                    const int hh = h + kh - 1;
                                    if (hh < 0 || hh >= H) continue;
                                    for (int kw = 0; kw < 3; ++kw) {
                                        const int ww = xw + kw - 1;
                                        if (ww < 0 || ww >= W) continue;
                                        acc += (double)xp[((size_t)dd * H + hh) * W + ww] * wp[(kd * 3 + kh) * 3 + kw];
                                    }
                                }
                            }
                        }
                        float v = (float)acc;
                        if (scale) v = v * scale[o] + shift[o];
                        if (relu && v < 0.f) v = 0.f;
                        const size_t idx = ((size_t)n * Cout + o) * S + ((size_t)d * H + h) * W + xw;
                        if (residual) v += residual[idx];
                        y[idx] = v;
                    }
    return 0;
}

/* MaxPool3d((1,2,2)) (stereo_network_old.py:156,165), NCDHW */
ORC_API int orc_maxpool_hw2(const float *x, float *y, int NC, int D, int H, int W)
{
    const int Ho = H / 2, Wo = W / 2;
    for (int i = 0; i < NC * D; ++i)
        for (int h = 0; h < Ho; ++h)
            for (int w = 0; w < Wo; ++w) {
                const float *p = x + ((size_t)i * H + 2 * h) * W + 2 * w;
                float m = p[0];
                if (p[1] > m) m = p[1];
                if (p[W] > m) m = p[W];
                if (p[W + 1] > m) m = p[W + 1];
                y[((size_t)i * Ho + h) * Wo + w] = m;
            }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * F2: dense photometric alignment (src/lib/dense_align/dense_align.py, box_3d.py).
 * Pinned by tests/golden/dense_align.npz, generated by EXECUTING the reference's align_parallel /
 * sample / enumeration_depth (oracle/gen_golden.py:gen_dense_align).  F.interpolate and F.grid_sample are
 * ATen ops (torch 2.11 in the image: grid_sample defaults to align_corners=False); restated from
 * ATen/native/UpSample.h (area_pixel_compute_source_index) and GridSampler.h (unnormalize + clip, bilinear).
 * --------------------------------------------------------------------------------------------- */
/* align_parallel's host preparation (dense_align.py:251-266): (img/255 - mean)/std, HWC->CHW, 2x bilinear up-sampling */
static void orc_up2_src(int dst, int size, int *i0, int *ip, float *l0, float *l1)
{
    float src = 0.5f * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    *i0 = (int)src;
    *ip = (*i0 < size - 1) ? 1 : 0;
    *l1 = src - (float)*i0;
    *l0 = 1.f - *l1;
}
ORC_API int orc_da_prep_u8(const unsigned char *img, const float *mean, const float *sd, float *out, int H, int W)
{
    const int H2 = 2 * H, W2 = 2 * W;
    float *norm = (float *)malloc(sizeof(float) * 3 * (size_t)H * W);
    if (!norm) return -1;
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < H * W; ++i) {
            const float v = (float)img[(size_t)i * 3 + c] / 255.f;
            norm[(size_t)c * H * W + i] = (v - mean[c]) / sd[c];
        }
    for (int c = 0; c < 3; ++c)
        for (int y = 0; y < H2; ++y) {
            int y0, yp; float h0, h1;
            orc_up2_src(y, H, &y0, &yp, &h0, &h1);
            for (int x = 0; x < W2; ++x) {
                int x0, xp; float w0, w1;
                orc_up2_src(x, W, &x0, &xp, &w0, &w1);
                const float *p = norm + (size_t)c * H * W + (size_t)y0 * W + x0;
                const float top = w0 * p[0] + w1 * p[xp];
                const float bot = w0 * p[(size_t)yp * W] + w1 * p[(size_t)yp * W + xp];
                out[((size_t)c * H2 + y) * W2 + x] = h0 * top + h1 * bot;
            }
        }
    free(norm);
    return 0;
}

static void orc_py_slice(int start, int stop, int len, int *s, int *e)
{
    if (start < 0) { start += len; if (start < 0) start = 0; } else if (start > len) start = len;
    if (stop < 0) { stop += len; if (stop < 0) stop = 0; } else if (stop > len) stop = len;
    *s = start; *e = stop;
}
static void orc_plane(const float *p1, const float *p2, const float *p3, float *pl)      /* box_3d.py:32-41 */
{
    const float a0 = p2[0] - p1[0], a1 = p2[1] - p1[1], a2 = p2[2] - p1[2];
    const float b0 = p3[0] - p1[0], b1 = p3[1] - p1[1], b2 = p3[2] - p1[2];
    const float n0 = a1 * b2 - a2 * b1, n1 = a2 * b0 - a0 * b2, n2 = a0 * b1 - a1 * b0;
    pl[0] = n0; pl[1] = n1; pl[2] = n2;
    pl[3] = -n0 * p1[0] - n1 * p1[1] - n2 * p1[2];
}
/* sample() (dense_align.py:14-70) for one RoI: writes up to cap valid (u, v, dz) triples in row-major order, returns the count */
ORC_API int orc_da_sample(const float *box, const float *borders, const float *poses, int rois, float f, float cx, float cy,
                          int f_h, int f_w, int cap, float *uvz, float *weight, int *count)
{
    static const int group[8][3] = {{0, 3, 4}, {2, 3, 4}, {1, 2, 4}, {0, 1, 4}, {0, 3, 5}, {2, 3, 5}, {1, 2, 5}, {0, 1, 5}};
    memset(uvz, 0, sizeof(float) * 3 * (size_t)rois * cap);
    memset(weight, 0, sizeof(float) * (size_t)rois * cap);
    for (int i = 0; i < rois; ++i) {
        const float *bx = box + 4 * i, *bd = borders + 2 * i, *ps = poses + 7 * i;
        int width = (int)((bd[1] - bd[0]) / 56.f), height = (int)((bx[3] - bx[1]) / 56.f);
        if (width < 1) width = 1;
        if (height < 1) height = 1;
        int r0, r1, c0, c1;
        orc_py_slice((int)((bx[1] + bx[3]) / 2.f + 0.5f), (int)(bx[3] - (bx[3] - bx[1]) * 0.1f + 0.5f), f_h, &r0, &r1);
        orc_py_slice((int)(bd[0] + 0.5f), (int)(bd[1] + 0.5f), f_w, &c0, &c1);
        /* Box3d.__init__ (box_3d.py:10-56) */
        const float cs = (float)cos((double)ps[6]), sn = (float)sin((double)ps[6]);
        const float R[9] = {cs, 0.f, sn, 0.f, 1.f, 0.f, -sn, 0.f, cs};
        const float hw = ps[3] / 2.f, hl = ps[5] / 2.f, hh = ps[4];
        const float Po[8][3] = {{-hw, 0.f, -hl}, {-hw, 0.f, hl}, {hw, 0.f, hl}, {hw, 0.f, -hl},
                                {-hw, -hh, -hl}, {-hw, -hh, hl}, {hw, -hh, hl}, {hw, -hh, -hl}};
        float Pc[8][3], planes[6][4];
        int nearest = 0;
        float nd = 100000000.f;
        for (int k = 0; k < 8; ++k) {
            for (int r = 0; r < 3; ++r)
                Pc[k][r] = ((R[3 * r] * Po[k][0] + R[3 * r + 1] * Po[k][1]) + R[3 * r + 2] * Po[k][2]) + ps[r];
            const float nrm = sqrtf((Pc[k][0] * Pc[k][0] + Pc[k][1] * Pc[k][1]) + Pc[k][2] * Pc[k][2]);
            if (nrm < nd) { nd = nrm; nearest = k; }
        }
        orc_plane(Pc[0], Pc[3], Pc[4], planes[0]);
        orc_plane(Pc[2], Pc[3], Pc[6], planes[1]);
        orc_plane(Pc[1], Pc[2], Pc[5], planes[2]);
        orc_plane(Pc[0], Pc[1], Pc[4], planes[3]);
        orc_plane(Pc[0], Pc[1], Pc[2], planes[4]);
        orc_plane(Pc[4], Pc[5], Pc[6], planes[5]);
        const float lo[3] = {-hw - 0.01f, -hh - 0.01f, -hl - 0.01f}, hi[3] = {hw + 0.01f, 0.f + 0.01f, hl + 0.01f};
        int n = 0;
        for (int r = r0; r < r1; r += height)
            for (int c = c0; c < c1; c += width) {
                const float u = (float)c, v = (float)r;
                const float nu = (u - cx) / f, nv = (v - cy) / f;
                int valid = 0;
                float z = 0.f;
                for (int k = 0; k < 3 && !valid; ++k) {                /* BoxRayInsec + mask_out_box (box_3d.py:58-102) */
                    const float *pl = planes[group[nearest][k]];
                    float t = (nu * pl[0] + nv * pl[1]) + 1.f * pl[2];
                    t = -(1.f / t) * pl[3];
                    const float i0 = nu * t - ps[0], i1 = nv * t - ps[1], i2 = 1.f * t - ps[2];
                    const float o0 = (R[0] * i0 + R[3] * i1) + R[6] * i2;
                    const float o1 = (R[1] * i0 + R[4] * i1) + R[7] * i2;
                    const float o2 = (R[2] * i0 + R[5] * i1) + R[8] * i2;
                    z = i2;
                    valid = o0 >= lo[0] && o1 >= lo[1] && o2 >= lo[2] && o0 <= hi[0] && o1 <= hi[1] && o2 <= hi[2];
                }
                if (valid) {
                    if (n < cap) {
                        float *o = uvz + ((size_t)i * cap + n) * 3;
                        o[0] = u; o[1] = v; o[2] = z;
                        weight[(size_t)i * cap + n] = 1.f;
                    }
                    ++n;
                }
            }
        count[i] = n;
    }
    return 0;
}

/* F.grid_sample(bilinear, padding_mode='border') of one position in a planar [3][H][W] image */
static float orc_gs_unnorm(float g, int size, int align)
{
    float x = align ? ((g + 1.f) / 2.f) * (float)(size - 1) : ((g + 1.f) * (float)size - 1.f) / 2.f;
    if (x < 0.f) x = 0.f;
    if (x > (float)(size - 1)) x = (float)(size - 1);
    return x;
}
static void orc_gs3(const float *im, int H, int W, float gx, float gy, int align, float *out)
{
    const float x = orc_gs_unnorm(gx, W, align), y = orc_gs_unnorm(gy, H, align);
    const float xw = floorf(x), yn = floorf(y);
    const float we = x - xw, ww = 1.f - we, wn = y - yn, ws = 1.f - wn;       /* GridSamplerKernel.cpp ApplyGridSample bilinear */
    const int ix = (int)xw, iy = (int)yn;
    const float nw = ws * ww, ne = ws * we, sw = wn * ww, se = wn * we;
    for (int c = 0; c < 3; ++c) {
        const float *p = im + (size_t)c * H * W;
        float acc = p[(size_t)iy * W + ix] * nw;
        if (ix + 1 < W) acc += p[(size_t)iy * W + ix + 1] * ne;
        if (iy + 1 < H) acc += p[(size_t)(iy + 1) * W + ix] * sw;
        if (ix + 1 < W && iy + 1 < H) acc += p[(size_t)(iy + 1) * W + ix + 1] * se;
        out[c] = acc;
    }
}
/* enumeration_depth (dense_align.py:175-237): err_sum [iters, rois], best_depth [rois], best_idx [rois] */
ORC_API int orc_da_enum(const float *imL, const float *imR, const float *uvz, const float *weight, const float *depth_enum,
                        float fb, int rois, int pixels, int iters, int H, int W, int align, float *err_sum, float *best_depth,
                        int *best_idx)
{
    const float half_w = (float)(((double)W - 1.0) / 2.0), half_h = (float)(((double)H - 1.0) / 2.0);
    for (int it = 0; it < iters; ++it)
        for (int r = 0; r < rois; ++r) {
            const float depth = depth_enum[(size_t)it * rois + r];
            const float dis = (1.f / depth) * fb;
            double s = 0;
            for (int p = 0; p < pixels; ++p) {
                const float *q = uvz + ((size_t)r * pixels + p) * 3;
                const float wgt = weight[(size_t)r * pixels + p];
                const float gy = (q[1] - half_h) / half_h;
                const float dd = 1.f / (q[2] / fb + 1.f / dis);
                float l[3], rr[3];
                orc_gs3(imL, H, W, (q[0] - half_w) / half_w, gy, align, l);
                orc_gs3(imR, H, W, ((q[0] - dd) - half_w) / half_w, gy, align, rr);
                for (int c = 0; c < 3; ++c) s += fabsf((l[c] - rr[c]) * wgt);
            }
            err_sum[(size_t)it * rois + r] = (float)s;
        }
    for (int r = 0; r < rois; ++r) {
        int bi = 0;
        for (int it = 1; it < iters; ++it)
            if (err_sum[(size_t)it * rois + r] < err_sum[(size_t)bi * rois + r]) bi = it;
        best_depth[r] = depth_enum[(size_t)bi * rois + r];
        best_idx[r] = bi;
    }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * F3: instance voxel volume of the stereo_network_new variant (models/networks/stereo_network_new.py).
 * get_voxel (:160-283) and the sampling of forward (:409-449).  Pinned by tests/golden/voxel_new.npz, generated by
 * EXECUTING the reference's get_voxel and stereo_network.forward (the voxel tensor is captured at the input of
 * self.pointNet).  F.grid_sample is an ATen op (bilinear, zeros padding, align_corners=False under torch 2.11).
 * --------------------------------------------------------------------------------------------- */
typedef struct { float x, y, z, depth; int b; } orc_vx_roi;
static orc_vx_roi orc_vx_roi_of(const float *lb, const float *rb, const float *p2, const float *fb, const float *trans_inv, int B)
{
    orc_vx_roi r;
    r.b = (int)lb[0];
    if (r.b < 0) r.b = 0;
    if (r.b > B - 1) r.b = B - 1;
    const float *ti = trans_inv + 6 * r.b, *P = p2 + 12 * r.b;
#define ORC_TX(x, y) (((x) * ti[0] + (y) * ti[1]) + ti[2])
#define ORC_TY(x, y) (((x) * ti[3] + (y) * ti[4]) + ti[5])
    const float cx = (ORC_TX(lb[1], lb[2]) + ORC_TX(lb[3], lb[4])) / 2.f;
    const float cy = (ORC_TY(lb[1], lb[2]) + ORC_TY(lb[3], lb[4])) / 2.f;
    const float cxr = (ORC_TX(rb[1], rb[2]) + ORC_TX(rb[3], rb[4])) / 2.f;
#undef ORC_TX
#undef ORC_TY
    r.depth = fb[r.b] / (cx - cxr);
    r.z = r.depth - P[11];
    r.x = ((cx * r.depth - P[3]) - P[2] * r.z) / P[0];
    r.y = ((cy * r.depth - P[7]) - P[6] * r.z) / P[5];
    return r;
}
static void orc_vx_project(const float *P, const float *tr, float X, float Y, float Z, float *uf, float *vf)
{
    const float a = ((X * P[0] + Y * P[1]) + Z * P[2]) + P[3];
    const float b = ((X * P[4] + Y * P[5]) + Z * P[6]) + P[7];
    const float w = ((X * P[8] + Y * P[9]) + Z * P[10]) + P[11];
    const float u = a / w, v = b / w, one = w / w;
    *uf = (u * tr[0] + v * tr[1]) + one * tr[2];
    *vf = (u * tr[3] + v * tr[4]) + one * tr[5];
}
static void orc_vx_voxel(const orc_vx_roi *r, int v, float *X, float *Y, float *Z)
{
    const int ix = v / 100, iy = (v / 10) % 10, iz = v % 10;
    *X = ((-2.5f + 0.5f * (float)ix) + 0.25f) + r->x;
    *Y = ((-2.5f + 0.5f * (float)iy) + 0.25f) + r->y;
    *Z = ((-5.f + (float)iz) + 0.5f) + r->z;
}
ORC_API int orc_voxel_coords(const float *left, const float *right, const float *p2, const float *p3, const float *fb,
                             const float *trans, const float *trans_inv, const float *depth_bins, int N, int B, int D,
                             int input_h, int input_w, float *norm3, float *valid3, float *normL, float *validL, float *normR,
                             float *validR, float *depth_ori)
{
    const float umax = (float)(input_w / 4.0 - 1.0), vmax = (float)(input_h / 4.0 - 1.0);
    for (int n = 0; n < N; ++n) {
        const float *lb = left + 5 * n;
        const orc_vx_roi r = orc_vx_roi_of(lb, right + 5 * n, p2, fb, trans_inv, B);
        depth_ori[n] = r.depth;
        float dmin = depth_bins[(size_t)n * D], dmax = dmin;
        for (int i = 1; i < D; ++i) {
            const float d = depth_bins[(size_t)n * D + i];
            if (d < dmin) dmin = d;
            if (d > dmax) dmax = d;
        }
        for (int v = 0; v < 1000; ++v) {
            float X, Y, Z, uf, vf, ufr, vfr;
            orc_vx_voxel(&r, v, &X, &Y, &Z);
            orc_vx_project(p2 + 12 * r.b, trans + 6 * r.b, X, Y, Z, &uf, &vf);
            orc_vx_project(p3 + 12 * r.b, trans + 6 * r.b, X, Y, Z, &ufr, &vfr);
            const size_t o = (size_t)n * 1000 + v;
            const float a = (uf - lb[1]) / (lb[3] - lb[1]) * 2.f - 1.f, b = (vf - lb[2]) / (lb[4] - lb[2]) * 2.f - 1.f;
            const float c = (Z - dmin) / (dmax - dmin) * 2.f - 1.f;
            norm3[3 * o] = a; norm3[3 * o + 1] = b; norm3[3 * o + 2] = c;
            valid3[o] = (a >= -1.f && a <= 1.f && b >= -1.f && b <= 1.f && c >= -1.f && c <= 1.f) ? 1.f : 0.f;
            const float lu = uf / umax * 2.f - 1.f, lv = vf / vmax * 2.f - 1.f, ru = ufr / umax * 2.f - 1.f, rv = vfr / vmax * 2.f - 1.f;
            normL[2 * o] = lu; normL[2 * o + 1] = lv; normR[2 * o] = ru; normR[2 * o + 1] = rv;
            validL[o] = (lu >= -1.f && lu <= 1.f && lv >= -1.f && lv <= 1.f) ? 1.f : 0.f;
            validR[o] = (ru >= -1.f && ru <= 1.f && rv >= -1.f && rv <= 1.f) ? 1.f : 0.f;
        }
    }
    return 0;
}
/* grid_sample(bilinear, zeros, align_corners) of all C channels of one NCHW image at normalised (gx, gy) */
static void orc_gs_zeros(const float *im, int C, int H, int W, float gx, float gy, int align, float *out, size_t ostride)
{
    const float x = align ? ((gx + 1.f) / 2.f) * (float)(W - 1) : ((gx + 1.f) * (float)W - 1.f) / 2.f;
    const float y = align ? ((gy + 1.f) / 2.f) * (float)(H - 1) : ((gy + 1.f) * (float)H - 1.f) / 2.f;
    const float xw = floorf(x), yn = floorf(y);
    const float we = x - xw, ww = 1.f - we, ws = y - yn, wn = 1.f - ws;
    const int ix = (int)xw, iy = (int)yn;
    const float w4[4] = {wn * ww, wn * we, ws * ww, ws * we};
    for (int c = 0; c < C; ++c) {
        const float *p = im + (size_t)c * H * W;
        float acc = 0.f;
        for (int k = 0; k < 4; ++k) {
            const int xx = ix + (k & 1), yy = iy + (k >> 1);
            if (xx >= 0 && xx < W && yy >= 0 && yy < H) acc += p[(size_t)yy * W + xx] * w4[k];
        }
        out[(size_t)c * ostride] = acc;
    }
}
/* forward (:409-449): voxel [N, 3C, 1000] = cat(L - R, L, R), samples of invalid voxels zeroed */
ORC_API int orc_voxel_volume(const float *featL, const float *featR, const float *left, const float *right, const float *p2,
                             const float *p3, const float *fb, const float *trans, const float *trans_inv, int N, int B, int C,
                             int H, int W, int input_h, int input_w, int align, float *voxel, float *depth_ori)
{
    const float umax = (float)(input_w / 4.0 - 1.0), vmax = (float)(input_h / 4.0 - 1.0);
    float *l = (float *)malloc(sizeof(float) * 2 * (size_t)C), *rr = l + C;
    if (!l) return -1;
    for (int n = 0; n < N; ++n) {
        const orc_vx_roi r = orc_vx_roi_of(left + 5 * n, right + 5 * n, p2, fb, trans_inv, B);
        depth_ori[n] = r.depth;
        for (int v = 0; v < 1000; ++v) {
            float X, Y, Z, uf, vf, ufr, vfr;
            orc_vx_voxel(&r, v, &X, &Y, &Z);
            orc_vx_project(p2 + 12 * r.b, trans + 6 * r.b, X, Y, Z, &uf, &vf);
            orc_vx_project(p3 + 12 * r.b, trans + 6 * r.b, X, Y, Z, &ufr, &vfr);
            const float lu = uf / umax * 2.f - 1.f, lv = vf / vmax * 2.f - 1.f, ru = ufr / umax * 2.f - 1.f, rv = vfr / vmax * 2.f - 1.f;
            const int okl = lu >= -1.f && lu <= 1.f && lv >= -1.f && lv <= 1.f, okr = ru >= -1.f && ru <= 1.f && rv >= -1.f && rv <= 1.f;
            for (int c = 0; c < C; ++c) l[c] = rr[c] = 0.f;
            if (okl) orc_gs_zeros(featL + (size_t)r.b * C * H * W, C, H, W, lu, lv, align, l, 1);
            if (okr) orc_gs_zeros(featR + (size_t)r.b * C * H * W, C, H, W, ru, rv, align, rr, 1);
            float *o = voxel + (size_t)n * 3 * C * 1000 + v;
            for (int c = 0; c < C; ++c) {
                o[(size_t)c * 1000] = l[c] - rr[c];
                o[(size_t)(C + c) * 1000] = l[c];
                o[(size_t)(2 * C + c) * 1000] = rr[c];
            }
        }
    }
    free(l);
    return 0;
}
