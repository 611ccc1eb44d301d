// vol_common.cuh -- sampling geometry shared by the instance cost-volume kernels (inst_costvol.cu, inst_costvol_cl.cu):
// get_proposal_shift (stereo_network_old.py:34-133) and torchvision's legacy RoIAlign sample positions
// (aligned=False, sampling_ratio 2; call site stereo_network_old.py:271, 372-373), one rounded float op per torch op.
#pragma once
#include "common.cuh"

namespace side {

struct AxisSample {
    int lo, hi;   // lo < 0  => sample outside the image (contributes 0)
    float l, h;   // fractional weight towards hi, and 1 - l
};

// Depth candidate i of RoI (l, r): stereo_network_old.py:53-77, one rounded float op per torch op.
__device__ __forceinline__ void proposal_for(const float *__restrict__ l, const float *__restrict__ r, float fb,
                                             int i, int D, float x_clamp, float &dbin, float &lx1, float &lx2,
                                             float &rx1, float &rx2, float &y1, float &y2)
{
    const float xmin = fminf(l[1], r[1]);
    const float xmax = fmaxf(l[3], r[3]);
    y1 = fminf(l[2], r[2]);
    y2 = fmaxf(l[4], r[4]);
    float t = __fsub_rn(xmax, xmin);
    t = __fmul_rn(t, 0.9f);
    t = __fmul_rn(t, 4.0f);
    float dmin = __fdiv_rn(fb, t);
    dmin = fminf(fmaxf(dmin, 1.0f), 87.0f);
    const float rate = (float)((double)i / (double)(D - 1));
    const float u = __fmul_rn(__fsub_rn(87.0f, dmin), rate);
    dbin = __fsub_rn(87.0f, u);
    const float disp = __fdiv_rn(__fdiv_rn(fb, dbin), 8.0f);
    lx1 = fminf(__fadd_rn(xmin, disp), x_clamp);
    lx2 = fminf(__fadd_rn(xmax, disp), x_clamp);
    rx1 = fmaxf(__fsub_rn(xmin, disp), 0.0f);
    rx2 = fmaxf(__fsub_rn(xmax, disp), 0.0f);
}

// One axis of torchvision's pre_calc_for_bilinear_interpolate (legacy aligned=False, sampling_ratio 2).
__device__ __forceinline__ AxisSample axis_sample(float start, float bin, int p, int i, int size)
{
    // (.. / 2) == (.. * 0.5f) bit for bit (power-of-two scaling); the multiply avoids the IEEE division subroutine
    float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                        __fmul_rn(__fmul_rn((float)i + 0.5f, bin), 0.5f));
    AxisSample s;
    if (c < -1.0f || c > (float)size) {
        s.lo = -1; s.hi = -1; s.l = 0.f; s.h = 0.f;
        return s;
    }
    if (c <= 0.f) c = 0.f;
    int lo = (int)c, hi;
    if (lo >= size - 1) {
        hi = lo = size - 1;
        c = (float)lo;
    } else {
        hi = lo + 1;
    }
    s.lo = lo; s.hi = hi;
    s.l = __fsub_rn(c, (float)lo);
    s.h = __fsub_rn(1.0f, s.l);
    return s;
}

struct VolParams {
    const float *featL, *featR, *left, *right, *fb;
    const uint8_t *valid;
    float *cost, *depth_bin, *xcross;
    const float *gcost;
    float *gfeatL, *gfeatR;
    int N, B, C, H, W, D, P;
    float x_clamp;
};

// Per-axis sample with the NHWC element offset folded in (row*W*C for y, col*C for x).  Samples outside the image
// keep offset 0 and get zero weights, which makes every tap product an exact zero without a branch
// (finite features assumed: 0 * inf would differ from the reference's hard zero).
struct AxisTap {
    int olo, ohi;
    float l, h;
};

__device__ __forceinline__ AxisTap to_tap(const AxisSample &s, int stride)
{
    AxisTap t;
    if (s.lo < 0) { t.olo = 0; t.ohi = 0; t.l = 0.f; t.h = 0.f; }
    else { t.olo = s.lo * stride; t.ohi = s.hi * stride; t.l = s.l; t.h = s.h; }
    return t;
}

// first / last integer cell touched by x-samples e0..e1 of a box (clamped into the image like axis_sample does)
__device__ __forceinline__ void sep_cells(float start, float bin, int e0, int e1, int W, int &c0, int &c1)
{
    const float a = __fadd_rn(__fadd_rn(start, __fmul_rn((float)(e0 >> 1), bin)),
                              __fmul_rn(__fmul_rn((float)(e0 & 1) + 0.5f, bin), 0.5f));
    const float b = __fadd_rn(__fadd_rn(start, __fmul_rn((float)(e1 >> 1), bin)),
                              __fmul_rn(__fmul_rn((float)(e1 & 1) + 0.5f, bin), 0.5f));
    c0 = min(max((int)floorf(fminf(fmaxf(a, -2.f), (float)W + 1.f)), 0), W - 1);
    c1 = min(min(max((int)floorf(fminf(fmaxf(b, -2.f), (float)W + 1.f)), 0), W - 1) + 1, W - 1);
}

}  // namespace side
