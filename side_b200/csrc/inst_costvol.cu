// inst_costvol.cu -- structure-aware instance depth cost volume.
//
// Reference behaviour replaced (all torch/torchvision ops, one launch each):
//   get_proposal_shift                      stereo_network_old.py:34-133
//   2*D RoIAlign launches + 3*D slice copies stereo_network_old.py:368-376   (RoIAlign((P,P),1,2), :271)
//   cosine gate x_cross                     stereo_network_old.py:197-203
//
// B200 design: one CTA per (RoI n, depth candidate d).  The CTA derives its two shifted RoIs from the
// boxes itself (no [D,N,5] proposal tensor round trip), builds separable sample tables once in shared
// memory (the y table is shared by the left and the right RoI), gathers L and R tiles [C,P,P] into shared
// memory, reduces the three gate sums with warp shuffles, and streams the gated [3C,P,P] slice to HBM with
// 16-byte evict-first stores -- the volume is written exactly once and never re-read by this kernel.
//
// Numerics: the sample coordinates, bilinear weights and the 4-tap / 4-sample sums are evaluated with
// explicitly rounded __f*_rn intrinsics in exactly the order torchvision's CPU kernel uses, so L, R and
// L-R are bit-identical to the oracle (oracle/side_oracle.c: roi_align_one); only the gate scalar differs
// (float32 reduction order), <= 1e-6 relative.
#include "vol_common.cuh"

namespace side {

// compiler-level scheduling barrier: keeps ptxas from hoisting the next batch of shared loads over this point
// (bounds the live registers of the separable slice loop; it emits no instruction)
#define SEP_SCHED_FENCE() asm volatile("" ::: "memory")

// sum over the 2x2 sample grid of one bin, torchvision order (iy outer, ix inner), then / 4
__device__ __forceinline__ float roi_bin(const float *__restrict__ im, int W, const AxisSample *__restrict__ ys,
                                         const AxisSample *__restrict__ xs)
{
    float acc = 0.f;
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
        const AxisSample y = ys[iy];
#pragma unroll
        for (int ix = 0; ix < 2; ++ix) {
            const AxisSample x = xs[ix];
            float val = 0.f;
            if (y.lo >= 0 && x.lo >= 0) {
                const float w1 = __fmul_rn(y.h, x.h), w2 = __fmul_rn(y.h, x.l);
                const float w3 = __fmul_rn(y.l, x.h), w4 = __fmul_rn(y.l, x.l);
                const float v1 = __ldg(im + y.lo * W + x.lo), v2 = __ldg(im + y.lo * W + x.hi);
                const float v3 = __ldg(im + y.hi * W + x.lo), v4 = __ldg(im + y.hi * W + x.hi);
                val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)),
                                __fmul_rn(w4, v4));
            }
            acc = __fadd_rn(acc, val);
        }
    }
    return __fdiv_rn(acc, 4.0f);
}

constexpr int kVolThreads = 512;
constexpr int kNhwcThreads = 1024;   // 60 regs/thread: one CTA of 32 warps per SM next to its 133 KB of tiles

// block-wide sum of up to 4 values; result broadcast to all threads.  red: >= 4*32 floats of smem
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float *red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) red[k * 32 + wid] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        float t = lane < nw ? red[k * 32 + lane] : 0.f;
        v[k] = warp_sum(t);
    }
}

// shared-memory carve-up: [tables 3*2P AxisSample][red 4*32 floats][Ls C*PP][Rs C*PP]
__device__ __forceinline__ void build_tables(const VolParams &p, int n, int d, AxisSample *ytab, AxisSample *xl,
                                             AxisSample *xr, int &b, float &dbin)
{
    const float *l = p.left + (size_t)n * 5, *r = p.right + (size_t)n * 5;
    b = min(max((int)l[0], 0), p.B - 1);
    float lx1, lx2, rx1, rx2, y1, y2;
    proposal_for(l, r, p.fb[b], d, p.D, p.x_clamp, dbin, lx1, lx2, rx1, rx2, y1, y2);
    const int t = threadIdx.x;
    const int P2 = 2 * p.P;
    if (t < 3 * P2) {
        const int which = t / P2, s = t % P2;
        if (which == 0) {
            const float rh = fmaxf(__fsub_rn(y2, y1), 1.0f);
            ytab[s] = axis_sample(y1, __fdiv_rn(rh, (float)p.P), s >> 1, s & 1, p.H);
        } else if (which == 1) {
            const float rw = fmaxf(__fsub_rn(lx2, lx1), 1.0f);
            xl[s] = axis_sample(lx1, __fdiv_rn(rw, (float)p.P), s >> 1, s & 1, p.W);
        } else {
            const float rw = fmaxf(__fsub_rn(rx2, rx1), 1.0f);
            xr[s] = axis_sample(rx1, __fdiv_rn(rw, (float)p.P), s >> 1, s & 1, p.W);
        }
    }
}

// GATE: apply x_cross.  STAGE: L/R tiles fit in shared memory (single gather pass, vector stores).
template <bool GATE, bool STAGE>
__global__ void __launch_bounds__(kVolThreads) inst_costvol_fwd_kernel(VolParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P = p.P, PP = P * P, C = p.C, CPP = C * PP;
    AxisSample *ytab = reinterpret_cast<AxisSample *>(smem_raw);
    AxisSample *xl = ytab + 2 * P, *xr = xl + 2 * P;
    float *red = reinterpret_cast<float *>(xr + 2 * P);
    float *Ls = red + 4 * 32, *Rs = Ls + CPP;

    const int n = blockIdx.x / p.D, d = blockIdx.x % p.D;
    const size_t cs = (size_t)p.D * PP;                       // channel stride of cost
    float *out = p.cost + (size_t)n * 3 * C * cs + (size_t)d * PP;

    if (p.valid && !p.valid[n]) {                              // dropped RoI: zero slice
        for (int e = threadIdx.x; e < 3 * CPP; e += blockDim.x) st_cs(out + (size_t)(e / PP) * cs + e % PP, 0.f);
        if (threadIdx.x == 0) {
            p.depth_bin[(size_t)n * p.D + d] = 0.f;
            if (GATE && p.xcross) p.xcross[(size_t)n * p.D + d] = 0.f;
        }
        return;
    }

    int b;
    float dbin;
    build_tables(p, n, d, ytab, xl, xr, b, dbin);
    if (threadIdx.x == 0) p.depth_bin[(size_t)n * p.D + d] = dbin;
    __syncthreads();

    const float *fL = p.featL + (size_t)b * C * p.H * p.W;
    const float *fR = p.featR + (size_t)b * C * p.H * p.W;
    const int HW = p.H * p.W;

    float s[3] = {0.f, 0.f, 0.f};
    for (int e = threadIdx.x; e < CPP; e += blockDim.x) {
        const int c = e / PP, q = e - c * PP, ph = q / P, pw = q - ph * P;
        const float l = roi_bin(fL + (size_t)c * HW, p.W, ytab + 2 * ph, xl + 2 * pw);
        const float r = roi_bin(fR + (size_t)c * HW, p.W, ytab + 2 * ph, xr + 2 * pw);
        if (GATE) {
            s[0] = fmaf(l, l, s[0]);
            s[1] = fmaf(r, r, s[1]);
            s[2] = fmaf(l, r, s[2]);
        }
        if (STAGE) {
            Ls[e] = l;
            Rs[e] = r;
        } else if (!GATE) {
            st_cs(out + (size_t)c * cs + q, l);
            st_cs(out + (size_t)(C + c) * cs + q, r);
            st_cs(out + (size_t)(2 * C + c) * cs + q, __fsub_rn(l, r));
        }
    }
    float g = 1.0f;
    if (GATE) {
        block_sum<3>(s, red);
        const float den = fmaxf(__fmul_rn(sqrtf(s[0]), sqrtf(s[1])), 0.01f);
        g = __fdiv_rn(s[2], den);
        if (threadIdx.x == 0 && p.xcross) p.xcross[(size_t)n * p.D + d] = g;
    }
    if (STAGE) {
        __syncthreads();
        if ((PP & 3) == 0) {
            const int PP4 = PP >> 2;
            for (int e4 = threadIdx.x; e4 < 3 * C * PP4; e4 += blockDim.x) {
                const int ch = e4 / PP4, q4 = e4 - ch * PP4;
                float4 v;
                if (ch < C) {
                    v = *reinterpret_cast<const float4 *>(Ls + ch * PP + 4 * q4);
                } else if (ch < 2 * C) {
                    v = *reinterpret_cast<const float4 *>(Rs + (ch - C) * PP + 4 * q4);
                } else {
                    const float4 a = *reinterpret_cast<const float4 *>(Ls + (ch - 2 * C) * PP + 4 * q4);
                    const float4 bq = *reinterpret_cast<const float4 *>(Rs + (ch - 2 * C) * PP + 4 * q4);
                    v = make_float4(__fsub_rn(a.x, bq.x), __fsub_rn(a.y, bq.y), __fsub_rn(a.z, bq.z),
                                    __fsub_rn(a.w, bq.w));
                }
                if (GATE) {
                    v.x = __fmul_rn(v.x, g); v.y = __fmul_rn(v.y, g);
                    v.z = __fmul_rn(v.z, g); v.w = __fmul_rn(v.w, g);
                }
                st_cs(reinterpret_cast<float4 *>(out + (size_t)ch * cs + 4 * q4), v);
            }
        } else {
            for (int e = threadIdx.x; e < 3 * CPP; e += blockDim.x) {
                const int ch = e / PP, q = e - ch * PP;
                float v = ch < C ? Ls[e] : (ch < 2 * C ? Rs[e - CPP] : __fsub_rn(Ls[e - 2 * CPP], Rs[e - 2 * CPP]));
                if (GATE) v = __fmul_rn(v, g);
                st_cs(out + (size_t)ch * cs + q, v);
            }
        }
    } else if (GATE) {
        // tiles do not fit in shared memory: second gather pass (features are L2 resident)
        for (int e = threadIdx.x; e < CPP; e += blockDim.x) {
            const int c = e / PP, q = e - c * PP, ph = q / P, pw = q - ph * P;
            const float l = roi_bin(fL + (size_t)c * HW, p.W, ytab + 2 * ph, xl + 2 * pw);
            const float r = roi_bin(fR + (size_t)c * HW, p.W, ytab + 2 * ph, xr + 2 * pw);
            st_cs(out + (size_t)c * cs + q, __fmul_rn(l, g));
            st_cs(out + (size_t)(C + c) * cs + q, __fmul_rn(r, g));
            st_cs(out + (size_t)(2 * C + c) * cs + q, __fmul_rn(__fsub_rn(l, r), g));
        }
    }
}


// ------------------------------------------------------------------------------------------------
// channels-last gather path (default when a workspace is supplied).
// The 16 bilinear taps per output make the NCHW gather L1-issue bound (one 4-byte load per tap).  With the feature
// maps transposed once to NHWC (B*H*W*C, a 2 x 7.8 MB pass against a 604 MB volume at config #2) every tap is ONE
// 16-byte load that serves 4 channels, a warp reads 4 x 128 contiguous bytes per instruction, and taps shared by the
// 2x2 sub-samples of a bin (same integer cell) are loaded once.  Arithmetic is unchanged (same rounded ops, same
// order), so the result stays bit-identical to the oracle.
// ------------------------------------------------------------------------------------------------
// Per-axis sample with the NHWC element offset folded in (row*W*C for y, col*C for x).  Samples outside the image
// keep offset 0 and get zero weights, which makes every tap product an exact zero without a branch
// (finite features assumed: 0 * inf would differ from the reference's hard zero).
template <bool FAST>
__device__ __forceinline__ float tap_val(float w1, float w2, float w3, float w4, float a, float b, float c, float d)
{
    if (FAST) return fmaf(w4, d, fmaf(w3, c, fmaf(w2, b, w1 * a)));   // contracted, like nvcc compiles torchvision's CUDA kernel
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, a), __fmul_rn(w2, b)), __fmul_rn(w3, c)), __fmul_rn(w4, d));
}

// one bin, 4 consecutive channels; base points at channel 4*c4 of pixel (0,0) of image b (NHWC)
template <bool FAST>
__device__ __forceinline__ float4 roi_bin4(const float *__restrict__ base, const AxisTap *__restrict__ ys,
                                           const AxisTap *__restrict__ xs)
{
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
        const AxisTap y = ys[iy];
#pragma unroll
        for (int ix = 0; ix < 2; ++ix) {
            const AxisTap x = xs[ix];
            const float w1 = __fmul_rn(y.h, x.h), w2 = __fmul_rn(y.h, x.l), w3 = __fmul_rn(y.l, x.h), w4 = __fmul_rn(y.l, x.l);
            const float4 v1 = __ldg(reinterpret_cast<const float4 *>(base + (y.olo + x.olo)));
            const float4 v2 = __ldg(reinterpret_cast<const float4 *>(base + (y.olo + x.ohi)));
            const float4 v3 = __ldg(reinterpret_cast<const float4 *>(base + (y.ohi + x.olo)));
            const float4 v4 = __ldg(reinterpret_cast<const float4 *>(base + (y.ohi + x.ohi)));
            acc.x = __fadd_rn(acc.x, tap_val<FAST>(w1, w2, w3, w4, v1.x, v2.x, v3.x, v4.x));
            acc.y = __fadd_rn(acc.y, tap_val<FAST>(w1, w2, w3, w4, v1.y, v2.y, v3.y, v4.y));
            acc.z = __fadd_rn(acc.z, tap_val<FAST>(w1, w2, w3, w4, v1.z, v2.z, v3.z, v4.z));
            acc.w = __fadd_rn(acc.w, tap_val<FAST>(w1, w2, w3, w4, v1.w, v2.w, v3.w, v4.w));
        }
    }
    // x * 0.25 == x / 4 for every float (power-of-two scaling rounds identically)
    return make_float4(__fmul_rn(acc.x, 0.25f), __fmul_rn(acc.y, 0.25f), __fmul_rn(acc.z, 0.25f), __fmul_rn(acc.w, 0.25f));
}

// PT: compile-time RoI size (16 = the reference's roiSize) or 0 for the generic runtime-P version.
template <bool GATE, bool FAST, int PT>
__global__ void __launch_bounds__(kNhwcThreads) inst_costvol_fwd_nhwc_kernel(VolParams p, const float *__restrict__ nhwcL,
                                                                            const float *__restrict__ nhwcR)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P = PT > 0 ? PT : p.P, PP = P * P, C = p.C, PPs = PP + 1;
    AxisTap *ytab = reinterpret_cast<AxisTap *>(smem_raw);
    AxisTap *xl = ytab + 2 * P, *xr = xl + 2 * P;
    float *red = reinterpret_cast<float *>(xr + 2 * P);
    float *Ls = red + 4 * 32, *Rs = Ls + (size_t)C * PPs;

    const int n = blockIdx.x / p.D, d = blockIdx.x - n * p.D;
    const size_t cs = (size_t)p.D * PP;
    float *out = p.cost + (size_t)n * 3 * C * cs + (size_t)d * PP;
    const int CPP = C * PP;

    if (p.valid && !p.valid[n]) {
        for (int e = threadIdx.x; e < 3 * CPP; e += blockDim.x) st_cs(out + (size_t)(e / PP) * cs + e % PP, 0.f);
        if (threadIdx.x == 0) {
            p.depth_bin[(size_t)n * p.D + d] = 0.f;
            if (GATE && p.xcross) p.xcross[(size_t)n * p.D + d] = 0.f;
        }
        return;
    }
    // sample tables (same arithmetic as the NCHW kernel / the oracle), converted to NHWC element offsets
    int b;
    {
        const float *l = p.left + (size_t)n * 5, *r = p.right + (size_t)n * 5;
        b = min(max((int)l[0], 0), p.B - 1);
        float dbin, lx1, lx2, rx1, rx2, y1, y2;
        proposal_for(l, r, p.fb[b], d, p.D, p.x_clamp, dbin, lx1, lx2, rx1, rx2, y1, y2);
        if (threadIdx.x == 0) p.depth_bin[(size_t)n * p.D + d] = dbin;
        const int t = threadIdx.x, P2 = 2 * P;
        if (t < 3 * P2) {
            const int which = t / P2, s = t - which * P2;
            if (which == 0) {
                const float rh = fmaxf(__fsub_rn(y2, y1), 1.0f);
                ytab[s] = to_tap(axis_sample(y1, __fdiv_rn(rh, (float)P), s >> 1, s & 1, p.H), p.W * C);
            } else if (which == 1) {
                const float rw = fmaxf(__fsub_rn(lx2, lx1), 1.0f);
                xl[s] = to_tap(axis_sample(lx1, __fdiv_rn(rw, (float)P), s >> 1, s & 1, p.W), C);
            } else {
                const float rw = fmaxf(__fsub_rn(rx2, rx1), 1.0f);
                xr[s] = to_tap(axis_sample(rx1, __fdiv_rn(rw, (float)P), s >> 1, s & 1, p.W), C);
            }
        }
    }
    __syncthreads();

    const float *fL = nhwcL + (size_t)b * p.H * p.W * C;
    const float *fR = nhwcR + (size_t)b * p.H * p.W * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int c4l = lane & 7, ql = lane >> 3;
    const int nc4 = C >> 2, nqblk = PP >> 2;

    float s[3] = {0.f, 0.f, 0.f};
    for (int c4 = c4l; c4 < nc4; c4 += 8) {
        const float *bl = fL + 4 * c4, *br = fR + 4 * c4;
        float *lp0 = Ls + (size_t)(4 * c4) * PPs, *rp0 = Rs + (size_t)(4 * c4) * PPs;
        for (int qblk = warp; qblk < nqblk; qblk += nwarps) {
            const int q = qblk * 4 + ql;
            const int ph = q / P, pw = q - ph * P;
            const float4 l = roi_bin4<FAST>(bl, ytab + 2 * ph, xl + 2 * pw);
            const float4 r = roi_bin4<FAST>(br, ytab + 2 * ph, xr + 2 * pw);
            if (GATE) {
                s[0] = fmaf(l.x, l.x, fmaf(l.y, l.y, fmaf(l.z, l.z, fmaf(l.w, l.w, s[0]))));
                s[1] = fmaf(r.x, r.x, fmaf(r.y, r.y, fmaf(r.z, r.z, fmaf(r.w, r.w, s[1]))));
                s[2] = fmaf(l.x, r.x, fmaf(l.y, r.y, fmaf(l.z, r.z, fmaf(l.w, r.w, s[2]))));
            }
            float *lp = lp0 + q, *rp = rp0 + q;
            lp[0] = l.x; lp[PPs] = l.y; lp[2 * PPs] = l.z; lp[3 * PPs] = l.w;
            rp[0] = r.x; rp[PPs] = r.y; rp[2 * PPs] = r.z; rp[3 * PPs] = r.w;
        }
    }
    float g = 1.0f;
    if (GATE) {
        block_sum<3>(s, red);
        const float den = fmaxf(__fmul_rn(sqrtf(s[0]), sqrtf(s[1])), 0.01f);
        g = __fdiv_rn(s[2], den);
        if (threadIdx.x == 0 && p.xcross) p.xcross[(size_t)n * p.D + d] = g;
    }
    __syncthreads();
    // write-out: thread owns column q = tid % PP (when blockDim % PP == 0) and strides over channels
    if (PT > 0 && (kNhwcThreads % (PT * PT)) == 0) {
        const int q = threadIdx.x % PP, c0 = threadIdx.x / PP, cstep = kNhwcThreads / PP;
        const float *lq = Ls + q, *rq = Rs + q;
        float *oq = out + q;
        for (int c = c0; c < C; c += cstep) {
            const float lv = lq[(size_t)c * PPs], rv = rq[(size_t)c * PPs];
            float dv = __fsub_rn(lv, rv), l2 = lv, r2 = rv;
            if (GATE) { l2 = __fmul_rn(lv, g); r2 = __fmul_rn(rv, g); dv = __fmul_rn(dv, g); }
            st_cs(oq + (size_t)c * cs, l2);
            st_cs(oq + (size_t)(C + c) * cs, r2);
            st_cs(oq + (size_t)(2 * C + c) * cs, dv);
        }
    } else {
        for (int e = threadIdx.x; e < CPP; e += blockDim.x) {
            const int c = e / PP, q = e - c * PP;
            const float lv = Ls[(size_t)c * PPs + q], rv = Rs[(size_t)c * PPs + q];
            float dv = __fsub_rn(lv, rv), l2 = lv, r2 = rv;
            if (GATE) { l2 = __fmul_rn(lv, g); r2 = __fmul_rn(rv, g); dv = __fmul_rn(dv, g); }
            st_cs(out + (size_t)c * cs + q, l2);
            st_cs(out + (size_t)(C + c) * cs + q, r2);
            st_cs(out + (size_t)(2 * C + c) * cs + q, dv);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Separable fast path (SIDE_VOL_SEPARABLE).
// The RoIAlign of depth candidate d samples the SAME 32 feature rows for every d of a RoI (the candidates only shift
// the box along x, stereo_network_old.py:70-77), so the y half of the bilinear interpolation and the sum over the two
// y sub-samples of a bin are shared by all D slices.  One CTA owns (RoI n, 8 channels, 8 of the 16 bin rows: ROWS below): it builds
//     U[side][ph][x][c] = sum_{iy} ( hy * f[ylo][x][c] + ly * f[yhi][x][c] )          (once per group of slices)
// in shared memory for the window of columns the group's shifted boxes touch (read straight from the NCHW features,
// 16-byte vectors along x, a column interpolated once and scattered to the quads' sub-windows), and every output bin is then 4 taps
//     bin = sum_{ix} ( 0.25*hx * U[ph][xlo] + 0.25*lx * U[ph][xhi] )
// instead of 16 global taps and 33 rounded operations.  One warp per slice: lane = (channel pair, row of a pair, 4
// consecutive bins), 8-byte conflict-free shared loads; the 32 x-sample table entries of the slice are computed one per lane and exchanged by shuffles; a quarter-warp's
// 16-byte evict-first stores cover 128 contiguous bytes.  The gate statistics (sum L^2, sum R^2, sum L*R) are
// accumulated in the same pass as per-(n, d, channel-chunk) partials, so the fused network path needs ONE pass over the
// volume (the scalar gate is applied by the consumer, side_ncdhw_to_cl_split); SIDE_VOL_GATE on this path runs a
// statistics pass first.
// Numerics: same sample positions, validity rules and weights as the exact kernels; the products are re-associated
// (y before x), so values differ from torchvision's order by a few ulp (<= 1e-6 relative, tests allow 1e-5); L-R is
// still computed from the kernel's own L and R, channel placement is exact.
// ------------------------------------------------------------------------------------------------
constexpr int kSepCC = 8;             // channels per CTA
constexpr int kSepThreads = 512;
constexpr int kSepXq = 46;            // window columns per (side, lane-quad)
constexpr int kSepMaxD = 256;
// U[side][ph][cell][8 channels]: cell = 4 * (x - win[quad]) + quad, i.e. the four lane-quads of a warp (bins 0-3, 4-7,
// 8-11, 12-15) read from four interleaved sub-windows whose 32-byte cells sit in different bank octets BY CONSTRUCTION;
// the row stride is = 4 (mod 32) floats so the two rows a warp instruction touches are 4 banks apart: every shared load
// of the slice loop is conflict-free whatever the box geometry.
constexpr int kSepRowF = 4 * (kSepXq + 2) * kSepCC + 4;     // floats per ph row (1540): kSepXq real cells + 2 zero cells
constexpr int kSepUSide = 16 * kSepRowF;                     // floats per side (ROWS = 16)

// shared-memory load through a 32-bit shared-space address + immediate byte offset
template <int IMM>
__device__ __forceinline__ float2 sep_lds2(uint32_t addr)
{
    float2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "n"(IMM));
    return v;
}

// ROWS = 16: one CTA of 512 threads per SM owns all 16 bin rows of (RoI, 8 channels).  ROWS = 8: the bin rows are split over
// blockIdx.z = 0, 1; a CTA has 256 threads and half the window (98 KB), so TWO CTAs share an SM and the window build of one
// (global loads, no stores) runs under the slice phase (stores) of the other -- the build is NOT repeated, each CTA builds
// only its own rows.
template <bool WRITE, bool STATS, bool APPLY, int ROWS>
__global__ void __launch_bounds__(32 * ROWS, 16 / ROWS) inst_costvol_sep_kernel(VolParams p, float *__restrict__ partial)
{
    constexpr int kThreads = 32 * ROWS, kUSide = ROWS * kSepRowF, kZ = 16 / ROWS;
    extern __shared__ __align__(16) float U[];              // [2][ROWS][kSepRowF]
    __shared__ AxisTap ytab[32];
    __shared__ float4 geo[kSepMaxD];                         // lx1, bin_w(left), rx1, bin_w(right) per slice
    __shared__ short2 cellbuf[kSepMaxD][8];                  // (first, last) cell per slice and (side, quad)
    __shared__ int g_d1, g_win[8], g_wr[8];                  // current slice group: last slice, window start / width
    __shared__ int s_slow;

    // the row halves of a (RoI, chunk) are neighbours in dispatch order: they run at the same time and their 512-byte halves of
    // every 1 KB output plane meet in L2 before they are written back
    const int n = blockIdx.y, chunk = blockIdx.x / kZ, zhalf = blockIdx.x % kZ, nchunk = gridDim.x / kZ, c0 = chunk * kSepCC;
    const int ph0 = zhalf * ROWS;                            // first bin row of this CTA
    constexpr int kUnits = ROWS / 4;                         // groups of 4 bin rows
    const int nslot = nchunk * kZ, slot0 = chunk * kZ + zhalf;             // statistics partials: one slot per CTA of a RoI
    const int C = p.C, D = p.D, W = p.W;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kThreads >> 5;
    const size_t cs = (size_t)D * 256;
    float *outn = p.cost + ((size_t)n * 3 * C + c0) * cs;

    if (p.valid && !p.valid[n]) {
        if (WRITE) {
            for (int i = tid; i < 3 * kSepCC * D * 4 * ROWS; i += kThreads) {
                const int q4 = i % (4 * ROWS), rest = i / (4 * ROWS), d = rest % D, ch = rest / D;     // ch = which * 8 + cc
                st_cs(reinterpret_cast<float4 *>(outn + ((size_t)(ch >> 3) * C + (ch & 7)) * cs + (size_t)d * 256 + ph0 * 16) + q4,
                      make_float4(0.f, 0.f, 0.f, 0.f));
            }
            if (slot0 == 0)
                for (int d = tid; d < D; d += kThreads) p.depth_bin[(size_t)n * D + d] = 0.f;
        }
        if (STATS)
            for (int d = tid; d < D; d += kThreads)
                *reinterpret_cast<float4 *>(partial + (((size_t)n * D + d) * nslot + slot0) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }

    const float *lb = p.left + (size_t)n * 5, *rb = p.right + (size_t)n * 5;
    const int b = min(max((int)lb[0], 0), p.B - 1);
    const float fb = p.fb[b];
    if (tid == 0) s_slow = 0;
    for (int d = tid; d < D; d += kThreads) {
        float dbin, lx1, lx2, rx1, rx2, y1, y2;
        proposal_for(lb, rb, fb, d, D, p.x_clamp, dbin, lx1, lx2, rx1, rx2, y1, y2);
        const float rwl = fmaxf(__fsub_rn(lx2, lx1), 1.0f), rwr = fmaxf(__fsub_rn(rx2, rx1), 1.0f);
        geo[d] = make_float4(lx1, __fmul_rn(rwl, 0.0625f), rx1, __fmul_rn(rwr, 0.0625f));   // == rw / 16 exactly
        if (WRITE && slot0 == 0) p.depth_bin[(size_t)n * D + d] = dbin;
    }
    if (tid < 32) {
        float dbin, lx1, lx2, rx1, rx2, y1, y2;
        proposal_for(lb, rb, fb, 0, D, p.x_clamp, dbin, lx1, lx2, rx1, rx2, y1, y2);
        const float rh = fmaxf(__fsub_rn(y2, y1), 1.0f);
        ytab[tid] = to_tap(axis_sample(y1, __fmul_rn(rh, 0.0625f), tid >> 1, tid & 1, p.H), W);      // NCHW row offsets
    }
    __syncthreads();
    // cells touched by each lane-quad (x-samples 8q .. 8q+7) of each slice, both sides
    for (int i = tid; i < D * 8; i += kThreads) {
        const int d = i >> 3, sq = i & 7, q = sq & 3;
        const float4 g4 = geo[d];
        int a0, a1;
        sep_cells(sq < 4 ? g4.x : g4.z, sq < 4 ? g4.y : g4.w, 8 * q, 8 * q + 7, W, a0, a1);
        cellbuf[d][sq] = make_short2((short)a0, (short)a1);
        if (a1 - a0 + 1 > kSepXq) s_slow = 1;               // one slice alone overflows a sub-window: box > ~90 columns wide
    }
    __syncthreads();

    if (s_slow) {
        const size_t plane = (size_t)p.H * W;
        const float *fLb = p.featL + ((size_t)b * C + c0) * plane;
        const float *fRb = p.featR + ((size_t)b * C + c0) * plane;
        // Very wide box (> ~90 feature columns, or garbage): the same formula evaluated straight from global memory,
        // one slice at a time.  Results are identical to the fast path's.
        float *red = U;
        for (int d = 0; d < D; ++d) {
            const float4 g4 = geo[d];
            float sv[4] = {0.f, 0.f, 0.f, 0.f};
            float g = 1.0f;
            if (APPLY) {
                float t0 = 0.f, t1 = 0.f, t2 = 0.f;
                for (int j = 0; j < nslot; ++j) {
                    const float4 ps = *reinterpret_cast<const float4 *>(partial + (((size_t)n * D + d) * nslot + j) * 4);
                    t0 += ps.x; t1 += ps.y; t2 += ps.z;
                }
                g = __fdiv_rn(t2, fmaxf(__fmul_rn(sqrtf(t0), sqrtf(t1)), 0.01f));
                if (slot0 == 0 && tid == 0 && p.xcross) p.xcross[(size_t)n * D + d] = g;
            }
            for (int idx = tid; idx < 16 * ROWS * kSepCC; idx += kThreads) {
                const int c1 = idx & 7, q = (idx >> 3) + ph0 * 16, ph = q >> 4, pw = q & 15;
                const AxisTap t0 = ytab[2 * ph], t1 = ytab[2 * ph + 1];
                float lr[2];
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    const float *f = (side ? fRb : fLb) + (size_t)c1 * plane;
                    float t = 0.f;
#pragma unroll
                    for (int ix = 0; ix < 2; ++ix) {
                        const AxisSample sx = axis_sample(side ? g4.z : g4.x, side ? g4.w : g4.y, pw, ix, W);
                        float ulo = 0.f, uhi = 0.f, wq = 0.f;
                        if (sx.lo >= 0) {
                            const float *a = f + sx.lo, *bq = f + sx.hi;
                            ulo = fmaf(t0.l, __ldg(a + t0.ohi), t0.h * __ldg(a + t0.olo)) + fmaf(t1.l, __ldg(a + t1.ohi), t1.h * __ldg(a + t1.olo));
                            uhi = fmaf(t0.l, __ldg(bq + t0.ohi), t0.h * __ldg(bq + t0.olo)) + fmaf(t1.l, __ldg(bq + t1.ohi), t1.h * __ldg(bq + t1.olo));
                            wq = 0.25f * sx.l;
                        }
                        t = ix == 0 ? (0.25f - wq) * ulo : fmaf(0.25f - wq, ulo, t);
                        t = fmaf(wq, uhi, t);
                    }
                    lr[side] = t;
                }
                if (STATS) {
                    sv[0] = fmaf(lr[0], lr[0], sv[0]); sv[1] = fmaf(lr[1], lr[1], sv[1]); sv[2] = fmaf(lr[0], lr[1], sv[2]);
                }
                if (WRITE) {
                    float *o = outn + (size_t)c1 * cs + (size_t)d * 256 + q;
                    float lv = lr[0], rv = lr[1], dv = __fsub_rn(lr[0], lr[1]);
                    if (APPLY) { lv = __fmul_rn(lv, g); rv = __fmul_rn(rv, g); dv = __fmul_rn(dv, g); }
                    st_cs(o, lv);
                    st_cs(o + (size_t)C * cs, rv);
                    st_cs(o + (size_t)2 * C * cs, dv);
                }
            }
            if (STATS) {
                block_sum<4>(sv, red);
                if (tid == 0)
                    *reinterpret_cast<float4 *>(partial + (((size_t)n * D + d) * nslot + slot0) * 4) = make_float4(sv[0], sv[1], sv[2], 0.f);
                __syncthreads();
            }
        }
        return;
    }

    // lane = (channel pair ccl, row-in-pair phsel, quad): a quarter-warp's 16-byte stores cover 128 contiguous bytes
    const int quad = lane & 3, phsel = (lane >> 2) & 1, ccl = lane >> 3;

    int d0 = 0;
    while (d0 < D) {
        // ---- slice group [d0, d1]: grow while all 8 sub-windows stay within kSepXq columns (warp 0, one lane per window) ----
        if (warp == 0) {
            const int sq = lane & 7;
            int lo = cellbuf[d0][sq].x, hi = cellbuf[d0][sq].y, dd = d0;
            while (dd + 1 < D) {
                const short2 c = cellbuf[dd + 1][sq];
                const int nlo = min(lo, (int)c.x), nhi = max(hi, (int)c.y);
                if (!__all_sync(0xffffffffu, nhi - nlo + 1 <= kSepXq)) break;
                lo = nlo; hi = nhi; ++dd;
            }
            if (lane < 8) { g_win[sq] = lo; g_wr[sq] = hi - lo + 1; }
            if (lane == 0) g_d1 = dd;
        }
        __syncthreads();
        const int d1 = g_d1;
        // ---- build U: local cells 0..wr-1 real, cells wr and wr+1 = zeros (invalid samples read (wr, wr+1); a sample
        //      clamped at the right border reads (W-1, W) with weight 0 on the second) ----
        //      The features are read as they are (NCHW, no staging copy): a task = (side, bin row, channel, 4 consecutive columns);
        //      a warp covers 8 channels x 4 bin rows of one column group, loads the (up to) four feature rows of its y samples as
        //      16-byte vectors ALONG x and scatters the four results into the sub-windows of the quads that contain the column
        //      (bank = 4 ph + 8 quad + c: at most 2-way conflicts), so a column is loaded and interpolated once, not once per quad.
        {
            int win[8], wr[8], xa[2], nx4 = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { win[k] = g_win[k]; wr[k] = g_wr[k]; }
#pragma unroll
            for (int sd = 0; sd < 2; ++sd) {
                int lo = win[4 * sd], hi = win[4 * sd] + wr[4 * sd];
#pragma unroll
                for (int k = 1; k < 4; ++k) { lo = min(lo, win[4 * sd + k]); hi = max(hi, win[4 * sd + k] + wr[4 * sd + k]); }
                xa[sd] = lo & ~3;
                nx4 = max(nx4, (hi - xa[sd] + 3) >> 2);
            }
            const size_t plane = (size_t)p.H * W;
            const size_t img = ((size_t)b * C + c0) * plane;
            const bool vec = (W & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.featL) | reinterpret_cast<uintptr_t>(p.featR)) & 15) == 0;
            const int ntask = 2 * ROWS * 8 * nx4;
            // branch-light body so that the unrolled iterations' loads are all in flight together: the build is latency-bound
#pragma unroll 4
            for (int it = tid; it < ntask; it += kThreads) {
                const int c = it & 7, ph = (it >> 3) & (ROWS - 1), r = it / (8 * ROWS), x4 = r % nx4, side = r / nx4;
                const int x = xa[side] + 4 * x4;
                const AxisTap t0 = ytab[2 * (ph0 + ph)], t1 = ytab[2 * (ph0 + ph) + 1];
                const float *fx = (side ? p.featR : p.featL) + img + (size_t)c * plane;
                float4 a0, a1, b0, b1;
                if (vec) {
                    const float *fv = fx + min(x, W - 4);
                    // the two y sub-samples of a bin usually fall in the same row pair or in adjacent ones: reuse the rows already
                    // loaded (unconditional loads of all four rows were measured: 10 % slower, the L1 pipe is the busy unit)
                    a0 = __ldg(reinterpret_cast<const float4 *>(fv + t0.olo));
                    a1 = __ldg(reinterpret_cast<const float4 *>(fv + t0.ohi));
                    b0 = a0; b1 = a1;
                    if (t1.olo != t0.olo || t1.ohi != t0.ohi) {
                        b0 = (t1.olo == t0.ohi) ? a1 : __ldg(reinterpret_cast<const float4 *>(fv + t1.olo));
                        b1 = __ldg(reinterpret_cast<const float4 *>(fv + t1.ohi));
                    }
                } else {
                    const int x0 = min(x, W - 1), x1 = min(x + 1, W - 1), x2 = min(x + 2, W - 1), x3 = min(x + 3, W - 1);
                    a0 = make_float4(__ldg(fx + t0.olo + x0), __ldg(fx + t0.olo + x1), __ldg(fx + t0.olo + x2), __ldg(fx + t0.olo + x3));
                    a1 = make_float4(__ldg(fx + t0.ohi + x0), __ldg(fx + t0.ohi + x1), __ldg(fx + t0.ohi + x2), __ldg(fx + t0.ohi + x3));
                    b0 = make_float4(__ldg(fx + t1.olo + x0), __ldg(fx + t1.olo + x1), __ldg(fx + t1.olo + x2), __ldg(fx + t1.olo + x3));
                    b1 = make_float4(__ldg(fx + t1.ohi + x0), __ldg(fx + t1.ohi + x1), __ldg(fx + t1.ohi + x2), __ldg(fx + t1.ohi + x3));
                }
                float u[4];
                u[0] = fmaf(t0.l, a1.x, t0.h * a0.x) + fmaf(t1.l, b1.x, t1.h * b0.x);
                u[1] = fmaf(t0.l, a1.y, t0.h * a0.y) + fmaf(t1.l, b1.y, t1.h * b0.y);
                u[2] = fmaf(t0.l, a1.z, t0.h * a0.z) + fmaf(t1.l, b1.z, t1.h * b0.z);
                u[3] = fmaf(t0.l, a1.w, t0.h * a0.w) + fmaf(t1.l, b1.w, t1.h * b0.w);
                float *urow = U + side * kUSide + (size_t)ph * kSepRowF + c;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xx = x + j;
                    const float v = xx <= W - 1 ? u[j] : 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cell = xx - (side ? win[4 + q] : win[q]);
                        if (cell >= 0 && cell < (side ? wr[4 + q] : wr[q])) urow[(4 * cell + q) * kSepCC] = v;
                    }
                }
            }
            // the two zero cells behind each sub-window
            for (int it = tid; it < 2 * 4 * ROWS * 2 * kSepCC; it += kThreads) {
                const int c = it & 7, z = (it >> 3) & 1, q = (it >> 4) & 3, ph = (it >> 6) & (ROWS - 1), side = it >> (6 + (ROWS == 16 ? 4 : 3));
                U[side * kUSide + (size_t)ph * kSepRowF + (4 * (g_wr[4 * side + q] + z) + q) * kSepCC + c] = 0.f;
            }
        }
        __syncthreads();
        // ---- one warp per slice (units of (slice, 4 bin rows) were measured: the repeated per-unit set-up costs more than the
        //      better balance of the last round returns) ----
        for (int d = d0 + warp; d <= d1; d += nwarps) {
            const float4 g4 = geo[d];
            // x-sample `lane` of this slice (it belongs to quad lane >> 3): byte offset of its first cell inside that quad's
            // sub-window + 0.25 * fractional weight
            uint32_t myoL, myoR;
            float mybL = 0.f, mybR = 0.f;
            {
                const int sq = lane >> 3;
                myoL = (uint32_t)((4 * g_wr[sq] + sq) * kSepCC * 4);         // invalid sample: the two zero cells
                myoR = (uint32_t)((4 * g_wr[4 + sq] + sq) * kSepCC * 4);
                const AxisSample sl = axis_sample(g4.x, g4.y, lane >> 1, lane & 1, W);
                if (sl.lo >= 0) {
                    myoL = (uint32_t)((4 * (sl.lo - g_win[sq]) + sq) * kSepCC * 4);
                    mybL = 0.25f * sl.l;
                }
                const AxisSample sr = axis_sample(g4.z, g4.w, lane >> 1, lane & 1, W);
                if (sr.lo >= 0) {
                    myoR = (uint32_t)((4 * (sr.lo - g_win[4 + sq]) + sq) * kSepCC * 4);
                    mybR = 0.25f * sr.l;
                }
            }
            // the second cell of a sample is always the next one of the sub-window (+128 bytes): hi = lo + 1, or lo == W-1
            // where the weight of the second cell is exactly 0 and the next cell holds zeros
            uint32_t oL[4], oR[4];        // two samples' offsets per register
            float bL[8], bR[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int src = (quad << 3) + j;
                const uint32_t tl = __shfl_sync(0xffffffffu, myoL, src), tr = __shfl_sync(0xffffffffu, myoR, src);
                if (j & 1) { oL[j >> 1] |= tl << 16; oR[j >> 1] |= tr << 16; }
                else { oL[j >> 1] = tl; oR[j >> 1] = tr; }
                bL[j] = __shfl_sync(0xffffffffu, mybL, src);
                bR[j] = __shfl_sync(0xffffffffu, mybR, src);
            }
            float g = 1.0f;
            if (APPLY) {
                float t0 = 0.f, t1 = 0.f, t2 = 0.f;
                for (int j = lane; j < nslot; j += 32) {
                    const float4 ps = *reinterpret_cast<const float4 *>(partial + (((size_t)n * D + d) * nslot + j) * 4);
                    t0 += ps.x; t1 += ps.y; t2 += ps.z;
                }
                t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
                const float den = fmaxf(__fmul_rn(sqrtf(t0), sqrtf(t1)), 0.01f);
                g = __fdiv_rn(t2, den);
                if (slot0 == 0 && lane == 0 && p.xcross) p.xcross[(size_t)n * D + d] = g;
            }
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            constexpr int kRowBytes = kSepRowF * 4;
            constexpr int kR = kUSide * 4;                            // byte distance of the right-view table
            const size_t side_stride = (size_t)C * cs;                // L -> R -> L-R planes
            uint32_t uL = smem_u32(U + phsel * kSepRowF + 2 * ccl);   // this lane's channel pair (2 ccl, 2 ccl + 1)
            float *o = outn + (size_t)(2 * ccl) * cs + (size_t)d * 256 + (ph0 + phsel) * 16 + 4 * quad;
#pragma unroll 1
            for (int rp2 = 0; rp2 < kUnits; ++rp2, uL += 4 * kRowBytes, o += 64) {
#pragma unroll
                for (int rp = 0; rp < 2; ++rp) {                      // row pair within this iteration
                    float l0[4], l1[4], r0[4], r1[4];                 // [bin] for the two channels, left / right view
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t a0 = uL + (oL[j] & 0xffffu), a1 = uL + (oL[j] >> 16);
                        const uint32_t c0r = uL + (oR[j] & 0xffffu), c1r = uL + (oR[j] >> 16);
                        float2 p0, p1, p2, p3, q0, q1, q2, q3;
                        if (rp == 0) {
                            p0 = sep_lds2<0>(a0); p1 = sep_lds2<128>(a0); p2 = sep_lds2<0>(a1); p3 = sep_lds2<128>(a1);
                            q0 = sep_lds2<kR>(c0r); q1 = sep_lds2<kR + 128>(c0r); q2 = sep_lds2<kR>(c1r); q3 = sep_lds2<kR + 128>(c1r);
                        } else {
                            p0 = sep_lds2<2 * kRowBytes>(a0); p1 = sep_lds2<2 * kRowBytes + 128>(a0);
                            p2 = sep_lds2<2 * kRowBytes>(a1); p3 = sep_lds2<2 * kRowBytes + 128>(a1);
                            q0 = sep_lds2<kR + 2 * kRowBytes>(c0r); q1 = sep_lds2<kR + 2 * kRowBytes + 128>(c0r);
                            q2 = sep_lds2<kR + 2 * kRowBytes>(c1r); q3 = sep_lds2<kR + 2 * kRowBytes + 128>(c1r);
                        }
                        const float w0 = bL[2 * j], w1 = bL[2 * j + 1], wa0 = 0.25f - w0, wa1 = 0.25f - w1;
                        l0[j] = fmaf(w1, p3.x, fmaf(wa1, p2.x, fmaf(w0, p1.x, wa0 * p0.x)));
                        l1[j] = fmaf(w1, p3.y, fmaf(wa1, p2.y, fmaf(w0, p1.y, wa0 * p0.y)));
                        const float v0 = bR[2 * j], v1 = bR[2 * j + 1], va0 = 0.25f - v0, va1 = 0.25f - v1;
                        r0[j] = fmaf(v1, q3.x, fmaf(va1, q2.x, fmaf(v0, q1.x, va0 * q0.x)));
                        r1[j] = fmaf(v1, q3.y, fmaf(va1, q2.y, fmaf(v0, q1.y, va0 * q0.y)));
                    }
                    if (STATS) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            s0 = fmaf(l0[j], l0[j], fmaf(l1[j], l1[j], s0));
                            s1 = fmaf(r0[j], r0[j], fmaf(r1[j], r1[j], s1));
                            s2 = fmaf(l0[j], r0[j], fmaf(l1[j], r1[j], s2));
                        }
                    }
                    if (WRITE) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const float *l = k ? l1 : l0, *r = k ? r1 : r0;
                            float4 lv = make_float4(l[0], l[1], l[2], l[3]), rv = make_float4(r[0], r[1], r[2], r[3]);
                            float4 dv = make_float4(__fsub_rn(l[0], r[0]), __fsub_rn(l[1], r[1]), __fsub_rn(l[2], r[2]),
                                                    __fsub_rn(l[3], r[3]));
                            if (APPLY) {
                                lv.x = __fmul_rn(lv.x, g); lv.y = __fmul_rn(lv.y, g); lv.z = __fmul_rn(lv.z, g); lv.w = __fmul_rn(lv.w, g);
                                rv.x = __fmul_rn(rv.x, g); rv.y = __fmul_rn(rv.y, g); rv.z = __fmul_rn(rv.z, g); rv.w = __fmul_rn(rv.w, g);
                                dv.x = __fmul_rn(dv.x, g); dv.y = __fmul_rn(dv.y, g); dv.z = __fmul_rn(dv.z, g); dv.w = __fmul_rn(dv.w, g);
                            }
                            float *oo = o + (size_t)k * cs + rp * 32;
                            st_cs(reinterpret_cast<float4 *>(oo), lv);
                            st_cs(reinterpret_cast<float4 *>(oo + side_stride), rv);
                            st_cs(reinterpret_cast<float4 *>(oo + 2 * side_stride), dv);
                        }
                    }
                }
            }
            if (STATS) {
                s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
                if (lane == 0)
                    *reinterpret_cast<float4 *>(partial + (((size_t)n * D + d) * nslot + slot0) * 4) = make_float4(s0, s1, s2, 0.f);
            }
        }
        __syncthreads();
        d0 = d1 + 1;
    }
}

// xcross[n, d] from the per-chunk partial sums (fused network path: the gate is applied by the consumer)
__global__ void sep_finish_xcross_kernel(const float *__restrict__ partial, const uint8_t *__restrict__ valid,
                                         float *__restrict__ xcross, int ND, int D, int nchunk)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ND) return;
    if (valid && !valid[i / D]) { xcross[i] = 0.f; return; }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int j = 0; j < nchunk; ++j) {
        const float4 ps = *reinterpret_cast<const float4 *>(partial + ((size_t)i * nchunk + j) * 4);
        s0 += ps.x; s1 += ps.y; s2 += ps.z;
    }
    xcross[i] = __fdiv_rn(s2, fmaxf(__fmul_rn(sqrtf(s0), sqrtf(s1)), 0.01f));
}

// scatter g * w_i / 4 to the 4 corners of the 2x2 samples of one bin (torchvision roi_align backward)
__device__ __forceinline__ void roi_bin_scatter(float *__restrict__ gim, int W, const AxisSample *__restrict__ ys,
                                                const AxisSample *__restrict__ xs, float g)
{
    const float gq = g * 0.25f;
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
        const AxisSample y = ys[iy];
#pragma unroll
        for (int ix = 0; ix < 2; ++ix) {
            const AxisSample x = xs[ix];
            if (y.lo >= 0 && x.lo >= 0) {
                atomicAdd(gim + y.lo * W + x.lo, gq * (y.h * x.h));
                atomicAdd(gim + y.lo * W + x.hi, gq * (y.h * x.l));
                atomicAdd(gim + y.hi * W + x.lo, gq * (y.l * x.h));
                atomicAdd(gim + y.hi * W + x.hi, gq * (y.l * x.l));
            }
        }
    }
}

template <bool GATE, bool STAGE>
__global__ void __launch_bounds__(kVolThreads) inst_costvol_bwd_kernel(VolParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P = p.P, PP = P * P, C = p.C, CPP = C * PP;
    AxisSample *ytab = reinterpret_cast<AxisSample *>(smem_raw);
    AxisSample *xl = ytab + 2 * P, *xr = xl + 2 * P;
    float *red = reinterpret_cast<float *>(xr + 2 * P);
    float *Ls = red + 4 * 32, *Rs = Ls + CPP;

    const int n = blockIdx.x / p.D, d = blockIdx.x % p.D;
    if (p.valid && !p.valid[n]) return;
    const size_t cs = (size_t)p.D * PP;
    const float *G = p.gcost + (size_t)n * 3 * C * cs + (size_t)d * PP;

    int b;
    float dbin;
    build_tables(p, n, d, ytab, xl, xr, b, dbin);
    __syncthreads();
    const int HW = p.H * p.W;
    const float *fL = p.featL + (size_t)b * C * HW;
    const float *fR = p.featR + (size_t)b * C * HW;
    float *gL = p.gfeatL + (size_t)b * C * HW;
    float *gR = p.gfeatR + (size_t)b * C * HW;

    float xc = 1.f, gxc = 0.f, inv_den = 0.f, al = 0.f, ar = 0.f;
    if (GATE) {
        // recompute L, R and the gate statistics, plus g_xc = sum(G * raw)
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int e = threadIdx.x; e < CPP; e += blockDim.x) {
            const int c = e / PP, q = e - c * PP, ph = q / P, pw = q - ph * P;
            const float l = roi_bin(fL + (size_t)c * HW, p.W, ytab + 2 * ph, xl + 2 * pw);
            const float r = roi_bin(fR + (size_t)c * HW, p.W, ytab + 2 * ph, xr + 2 * pw);
            if (STAGE) { Ls[e] = l; Rs[e] = r; }
            s[0] = fmaf(l, l, s[0]);
            s[1] = fmaf(r, r, s[1]);
            s[2] = fmaf(l, r, s[2]);
            const float g0 = G[(size_t)c * cs + q], g1 = G[(size_t)(C + c) * cs + q], g2 = G[(size_t)(2 * C + c) * cs + q];
            s[3] += g0 * l + g1 * r + g2 * (l - r);
        }
        block_sum<4>(s, red);
        const float nl = sqrtf(s[0]), nr = sqrtf(s[1]);
        const float prod = nl * nr;
        const bool active = prod > 0.01f;                    // clamp(min=0.01) passes gradient only above the floor
        const float den = active ? prod : 0.01f;
        xc = s[2] / den;
        gxc = s[3];
        inv_den = 1.0f / den;
        if (active && nl > 0.f && nr > 0.f) {
            al = s[2] * nr / (den * den * nl);
            ar = s[2] * nl / (den * den * nr);
        }
        if (STAGE) __syncthreads();
    }
    for (int e = threadIdx.x; e < CPP; e += blockDim.x) {
        const int c = e / PP, q = e - c * PP, ph = q / P, pw = q - ph * P;
        const float g0 = G[(size_t)c * cs + q], g1 = G[(size_t)(C + c) * cs + q], g2 = G[(size_t)(2 * C + c) * cs + q];
        float gl = g0 + g2, gr = g1 - g2;
        if (GATE) {
            float l, r;
            if (STAGE) { l = Ls[e]; r = Rs[e]; }
            else {
                l = roi_bin(fL + (size_t)c * HW, p.W, ytab + 2 * ph, xl + 2 * pw);
                r = roi_bin(fR + (size_t)c * HW, p.W, ytab + 2 * ph, xr + 2 * pw);
            }
            gl = xc * gl + gxc * (r * inv_den - al * l);
            gr = xc * gr + gxc * (l * inv_den - ar * r);
        }
        roi_bin_scatter(gL + (size_t)c * HW, p.W, ytab + 2 * ph, xl + 2 * pw, gl);
        roi_bin_scatter(gR + (size_t)c * HW, p.W, ytab + 2 * ph, xr + 2 * pw, gr);
    }
}

// ------------------------------------------------------------------------------------------------
// Separable backward, GATHER form (ungated volume; the gate's backward is side_xcross_gate_bwd).
//
// The scalar kernel above mirrors torchvision's roi_align backward: 32 fp32 atomics per volume element, 1.6 G atomics for
// config #2.  The forward is separable (y interpolation shared by all D candidates, see inst_costvol_sep_kernel), so its
// adjoint is too:
//     gU[side][ph][x][c] = sum_d sum_j 0.25 * tri(clamp(s_dj, 0, W-1) - x) * G[side][d][ph][j >> 1][c]      (x pass)
//     gfeat[y][x][c]    += sum_s  wy_s(y) * gU[ph = s >> 1][x][c]                                             (y pass)
// where s_dj are the 32 x-sample positions of slice d (an arithmetic progression, so the samples touching one column are
// a contiguous, computable range), tri(t) = max(0, 1 - |t|) is the bilinear weight and G = (g_L + g_{L-R}, g_R - g_{L-R}).
// A CTA owns (RoI, 8 channels): 4 producer warps stream the three gradient planes of one slice after another into an
// 8-deep shared-memory ring (already combined per side, scaled by 1/4 and laid out [pw][ph][c]), 12 consumer warps own
// (column, 8 bin rows) x 8 channels each with the sums in REGISTERS -- no atomics in the x pass at all.  The y pass runs
// once per CTA from shared memory and issues one atomic per (row, column, channel) of the RoI's window, coalesced along
// x: about 10 M atomics for config #2 instead of 1.6 G.
// ------------------------------------------------------------------------------------------------
constexpr int kBgCC = 8;
constexpr int kBgThreads = 512, kBgProd = 128, kBgCons = 384;
constexpr int kBgStages = 8;
constexpr int kBgPwF = 4 * 36;                  // floats per bin column pw: 4 row-quads x (4 rows x 8 channels + 4 pad)
constexpr int kBgSideF = 16 * kBgPwF;           // floats per side
constexpr int kBgStageF = 2 * kBgSideF;         // floats per ring slot (18 KB)
constexpr int kBgCols = kBgCons / 2;            // window columns (both sides together) per pass: 192 (a consumer thread owns 8 bin rows of one)
constexpr int kBgMaxRows = 512;

__global__ void __launch_bounds__(kBgThreads, 1) inst_costvol_bwd_gather_kernel(VolParams p)
{
    extern __shared__ __align__(16) float ring[];         // kBgStages slots; reused as gU[ph][c][col] by the y pass
    __shared__ float4 geo[kSepMaxD];                       // lx1, bin_w(left), rx1, bin_w(right) per slice
    __shared__ float2 ginv[kSepMaxD];                      // 2 / bin_w per side
    __shared__ AxisSample ytab[32];
    __shared__ short2 yrange[kBgMaxRows];                  // first / last y sample touching each row of the window
    __shared__ __align__(8) uint64_t full_bar[kBgStages], empty_bar[kBgStages];
    __shared__ int s_win[4], s_rows[2];

    const int n = blockIdx.y, c0 = blockIdx.x * kBgCC;
    if (p.valid && !p.valid[n]) return;
    const int C = p.C, D = p.D, W = p.W, H = p.H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *lb = p.left + (size_t)n * 5, *rb = p.right + (size_t)n * 5;
    const int b = min(max((int)lb[0], 0), p.B - 1);
    const float fb = p.fb[b];
    if (tid == 0) { s_win[0] = W; s_win[1] = -1; s_win[2] = W; s_win[3] = -1; }
    if (tid < kBgStages) { mbar_init(&full_bar[tid], kBgProd / 32); mbar_init(&empty_bar[tid], kBgCons / 32); }
    mbar_fence_init();
    __syncthreads();
    for (int d = tid; d < D; d += kBgThreads) {
        float dbin, lx1, lx2, rx1, rx2, y1, y2;
        proposal_for(lb, rb, fb, d, D, p.x_clamp, dbin, lx1, lx2, rx1, rx2, y1, y2);
        const float bl = __fmul_rn(fmaxf(__fsub_rn(lx2, lx1), 1.0f), 0.0625f), br = __fmul_rn(fmaxf(__fsub_rn(rx2, rx1), 1.0f), 0.0625f);
        geo[d] = make_float4(lx1, bl, rx1, br);
        ginv[d] = make_float2(2.0f / bl, 2.0f / br);
        int a0, a1;
        sep_cells(lx1, bl, 0, 31, W, a0, a1);
        atomicMin(&s_win[0], a0); atomicMax(&s_win[1], a1);
        sep_cells(rx1, br, 0, 31, W, a0, a1);
        atomicMin(&s_win[2], a0); atomicMax(&s_win[3], a1);
    }
    if (tid < 32) {
        float dbin, lx1, lx2, rx1, rx2, y1, y2;
        proposal_for(lb, rb, fb, 0, D, p.x_clamp, dbin, lx1, lx2, rx1, rx2, y1, y2);
        const float rh = fmaxf(__fsub_rn(y2, y1), 1.0f);
        ytab[tid] = axis_sample(y1, __fmul_rn(rh, 0.0625f), tid >> 1, tid & 1, H);
    }
    __syncthreads();
    if (tid == 0) {
        int y0 = H, y1 = -1;
        for (int s = 0; s < 32; ++s)
            if (ytab[s].lo >= 0) { y0 = min(y0, ytab[s].lo); y1 = max(y1, ytab[s].hi); }
        s_rows[0] = y0; s_rows[1] = y1;
    }
    __syncthreads();
    const int ymin = s_rows[0], nrows = s_rows[1] - s_rows[0] + 1;
    const int wl0 = s_win[0], wr0 = s_win[2];
    const int wL = max(s_win[1] - s_win[0] + 1, 0), wR = max(s_win[3] - s_win[2] + 1, 0);
    const int T = wL + wR;
    if (nrows <= 0 || T <= 0) return;                       // no valid sample: zero gradient
    for (int yy = tid; yy < nrows; yy += kBgThreads) {
        int s0 = 32, s1 = -1;
        for (int s = 0; s < 32; ++s)
            if (ytab[s].lo >= 0 && (ytab[s].lo == ymin + yy || ytab[s].hi == ymin + yy)) { s0 = min(s0, s); s1 = max(s1, s); }
        yrange[yy] = make_short2((short)s0, (short)s1);
    }
    const int npass = (T + kBgCols - 1) / kBgCols;
    const size_t cs = (size_t)D * 256;
    float *gfL = p.gfeatL + ((size_t)b * C + c0) * H * W, *gfR = p.gfeatR + ((size_t)b * C + c0) * H * W;
    int it = 0;                                             // slices streamed so far (ring position), all passes

    for (int pass = 0; pass < npass; ++pass) {
        if (warp < kBgProd / 32) {
            // ================= producers: lane = (bin row ph fastest, group of 4 bin columns), 4 channels each =================
            const int ph = tid & 15, pwq = (tid >> 4) & 3, chalf = tid >> 6;
            const float *g0 = p.gcost + ((size_t)n * 3 * C + c0 + chalf * 4) * cs + ph * 16 + pwq * 4;
            const int soff = (ph >> 2) * 36 + (ph & 3) * 8 + chalf * 4 + pwq * 4 * kBgPwF;
            for (int d = 0; d < D; ++d, ++it) {
                const int stage = it % kBgStages;
                float a[4][4], bq[4][4], cq[4][4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const float *gp = g0 + (size_t)cc * cs + (size_t)d * 256;
                    const float4 va = __ldcs(reinterpret_cast<const float4 *>(gp));
                    const float4 vb = __ldcs(reinterpret_cast<const float4 *>(gp + (size_t)C * cs));
                    const float4 vc = __ldcs(reinterpret_cast<const float4 *>(gp + (size_t)2 * C * cs));
                    a[cc][0] = va.x; a[cc][1] = va.y; a[cc][2] = va.z; a[cc][3] = va.w;
                    bq[cc][0] = vb.x; bq[cc][1] = vb.y; bq[cc][2] = vb.z; bq[cc][3] = vb.w;
                    cq[cc][0] = vc.x; cq[cc][1] = vc.y; cq[cc][2] = vc.z; cq[cc][3] = vc.w;
                }
                mbar_wait(&empty_bar[stage], ((uint32_t)(it / kBgStages) & 1u) ^ 1u);
                float *st = ring + (size_t)stage * kBgStageF + soff;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    *reinterpret_cast<float4 *>(st + j * kBgPwF) =
                        make_float4(0.25f * (a[0][j] + cq[0][j]), 0.25f * (a[1][j] + cq[1][j]), 0.25f * (a[2][j] + cq[2][j]),
                                    0.25f * (a[3][j] + cq[3][j]));
                    *reinterpret_cast<float4 *>(st + kBgSideF + j * kBgPwF) =
                        make_float4(0.25f * (bq[0][j] - cq[0][j]), 0.25f * (bq[1][j] - cq[1][j]), 0.25f * (bq[2][j] - cq[2][j]),
                                    0.25f * (bq[3][j] - cq[3][j]));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[stage]);
            }
            __syncthreads();                                // (A) consumers done with the ring
            __syncthreads();                                // (B) gU written
        } else {
            // ================= consumers: unit = (window column, 8 bin rows) x 8 channels, sums in registers =================
            const int ct = tid - kBgProd;
            const int phh = ct & 1;
            const int col = pass * kBgCols + (ct >> 1);
            const int uside = col >= T ? -1 : (col >= wL ? 1 : 0);
            const int ux = uside == 1 ? wr0 + col - wL : wl0 + col;
            float acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
            const float wmax = (float)(W - 1), wlim = (float)W, xf = (float)ux;
            for (int d = 0; d < D; ++d, ++it) {
                const int stage = it % kBgStages;
                const float4 g4 = geo[d];
                const float2 gi = ginv[d];
                mbar_wait(&full_bar[stage], (uint32_t)(it / kBgStages) & 1u);
                if (uside >= 0) {
                    const float start = uside ? g4.z : g4.x, bin = uside ? g4.w : g4.y, inv_h = uside ? gi.y : gi.x;
                    // samples s_j = start + (j + 0.5) * bin / 2 that can touch column x lie in (x - 1, x + 1): one candidate of
                    // slack on each side, the weights below decide
                    const float r0 = (xf - 1.f - start) * inv_h - 0.5f, r1 = (xf + 1.f - start) * inv_h - 0.5f;
                    const int jlo = max(0, (int)floorf(fminf(fmaxf(r0, -4.f), 64.f)));
                    const int jhi = min(31, (int)ceilf(fminf(fmaxf(r1, -4.f), 64.f)));
                    const float *sb = ring + (size_t)stage * kBgStageF + (uside ? kBgSideF : 0) + phh * 72;
                    for (int pw = jlo >> 1; pw <= (jhi >> 1); ++pw) {
                        float w = 0.f;
#pragma unroll
                        for (int ix = 0; ix < 2; ++ix) {
                            const float s = __fadd_rn(__fadd_rn(start, __fmul_rn((float)pw, bin)),
                                                      __fmul_rn(__fmul_rn((float)ix + 0.5f, bin), 0.5f));
                            const float sc = fminf(fmaxf(s, 0.f), wmax);
                            const float t = 1.f - fabsf(sc - xf);
                            if (s >= -1.f && s <= wlim && t > 0.f) w += t;
                        }
                        if (!(w > 0.f)) continue;
                        const float *q = sb + pw * kBgPwF;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 v0 = *reinterpret_cast<const float4 *>(q + (i >> 2) * 36 + (i & 3) * 8);
                            const float4 v1 = *reinterpret_cast<const float4 *>(q + (i >> 2) * 36 + (i & 3) * 8 + 4);
                            acc[i][0] = fmaf(w, v0.x, acc[i][0]); acc[i][1] = fmaf(w, v0.y, acc[i][1]);
                            acc[i][2] = fmaf(w, v0.z, acc[i][2]); acc[i][3] = fmaf(w, v0.w, acc[i][3]);
                            acc[i][4] = fmaf(w, v1.x, acc[i][4]); acc[i][5] = fmaf(w, v1.y, acc[i][5]);
                            acc[i][6] = fmaf(w, v1.z, acc[i][6]); acc[i][7] = fmaf(w, v1.w, acc[i][7]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
            }
            __syncthreads();                                // (A)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int c = 0; c < 8; ++c) ring[((phh * 8 + i) * kBgCC + c) * kBgCols + (ct >> 1)] = acc[i][c];
            __syncthreads();                                // (B)
        }
        // ================= y pass: a warp per (channel, row) of the window, lanes along the columns =================
        const int ncols = min(kBgCols, T - pass * kBgCols);
        for (int r = warp; r < kBgCC * nrows; r += kBgThreads / 32) {
            const int c = r / nrows, yy = r - c * nrows, y = ymin + yy;
            const short2 yr = yrange[yy];
            float *rowL = gfL + ((size_t)c * H + y) * W, *rowR = gfR + ((size_t)c * H + y) * W;
            for (int colL = lane; colL < ncols; colL += 32) {
                float v = 0.f;
                for (int s = yr.x; s <= yr.y; ++s) {
                    const AxisSample ys = ytab[s];
                    const float wy = (ys.lo == y ? ys.h : 0.f) + (ys.hi == y ? ys.l : 0.f);
                    v = fmaf(wy, ring[((s >> 1) * kBgCC + c) * kBgCols + colL], v);
                }
                if (v != 0.f) {
                    const int col = pass * kBgCols + colL;
                    if (col >= wL) atomicAdd(rowR + wr0 + col - wL, v);
                    else atomicAdd(rowL + wl0 + col, v);
                }
            }
        }
        __syncthreads();                                    // (C) ring free for the next pass
    }
}

__global__ void proposal_shift_kernel(const float *__restrict__ left, const float *__restrict__ right,
                                      const float *__restrict__ fb, int N, int B, int D, float x_clamp,
                                      float *__restrict__ pro_left, float *__restrict__ pro_right,
                                      float *__restrict__ depth_bin)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * D) return;
    const int n = t / D, i = t % D;
    const float *l = left + (size_t)n * 5, *r = right + (size_t)n * 5;
    float dbin, lx1, lx2, rx1, rx2, y1, y2;
    proposal_for(l, r, fb[min(max((int)l[0], 0), B - 1)], i, D, x_clamp, dbin, lx1, lx2, rx1, rx2, y1, y2);
    depth_bin[(size_t)n * D + i] = dbin;
    float *pl = pro_left + ((size_t)i * N + n) * 5, *pr = pro_right + ((size_t)i * N + n) * 5;
    pl[0] = l[0]; pl[1] = lx1; pl[2] = y1; pl[3] = lx2; pl[4] = y2;
    pr[0] = l[0]; pr[1] = rx1; pr[2] = y1; pr[3] = rx2; pr[4] = y2;
}

// ------------------------------------------------------------------------------------------------
// stand-alone gate on a materialised volume (drop-in cost_volume.forward entry)
// ------------------------------------------------------------------------------------------------
// `out` may alias `cost` (side_xcross_gate_fwd's contract; the gated volume backward runs in place): neither is __restrict__.
// Every element is read by the thread that later overwrites it, after the block-wide reduction.
template <bool BWD>
__global__ void __launch_bounds__(256) xcross_gate_kernel(const float *cost, const float *__restrict__ gout, float *out,
                                                          float *__restrict__ xcross, int C, int D, int PP)
{
    __shared__ float red[4 * 32];
    const int n = blockIdx.x / D, d = blockIdx.x % D;
    const size_t cs = (size_t)D * PP;
    const size_t base = (size_t)n * 3 * C * cs + (size_t)d * PP;
    const float *cb = cost + base;
    const int CPP = C * PP;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int e = threadIdx.x; e < CPP; e += blockDim.x) {
        const int c = e / PP, q = e - c * PP;
        const float l = cb[(size_t)c * cs + q], r = cb[(size_t)(C + c) * cs + q];
        s[0] = fmaf(l, l, s[0]);
        s[1] = fmaf(r, r, s[1]);
        s[2] = fmaf(l, r, s[2]);
        if (BWD) {
            const float *gb = gout + base;
            s[3] += gb[(size_t)c * cs + q] * l + gb[(size_t)(C + c) * cs + q] * r +
                    gb[(size_t)(2 * C + c) * cs + q] * cb[(size_t)(2 * C + c) * cs + q];
        }
    }
    block_sum<4>(s, red);
    const float nl = sqrtf(s[0]), nr = sqrtf(s[1]);
    const float prod = __fmul_rn(nl, nr);
    const bool active = prod > 0.01f;
    const float den = active ? prod : 0.01f;
    const float xc = __fdiv_rn(s[2], den);
    if (!BWD) {
        if (threadIdx.x == 0 && xcross) xcross[(size_t)n * D + d] = xc;
        float *ob = out + base;
        for (int e = threadIdx.x; e < 3 * CPP; e += blockDim.x) {
            const int c = e / PP, q = e - c * PP;
            ob[(size_t)c * cs + q] = __fmul_rn(cb[(size_t)c * cs + q], xc);
        }
    } else {
        const float gxc = s[3], inv_den = 1.0f / den;
        float al = 0.f, ar = 0.f;
        if (active && nl > 0.f && nr > 0.f) {
            al = s[2] * nr / (den * den * nl);
            ar = s[2] * nl / (den * den * nr);
        }
        const float *gb = gout + base;
        float *ob = out + base;
        for (int e = threadIdx.x; e < CPP; e += blockDim.x) {
            const int c = e / PP, q = e - c * PP;
            const float l = cb[(size_t)c * cs + q], r = cb[(size_t)(C + c) * cs + q];
            ob[(size_t)c * cs + q] = xc * gb[(size_t)c * cs + q] + gxc * (r * inv_den - al * l);
            ob[(size_t)(C + c) * cs + q] = xc * gb[(size_t)(C + c) * cs + q] + gxc * (l * inv_den - ar * r);
            ob[(size_t)(2 * C + c) * cs + q] = xc * gb[(size_t)(2 * C + c) * cs + q];
        }
    }
}

static size_t vol_smem_bytes(int C, int P, bool stage)
{
    size_t b = sizeof(AxisSample) * 6 * P + sizeof(float) * 4 * 32;
    if (stage) b += sizeof(float) * 2 * (size_t)C * P * P;
    return b;
}

static int check_vol_args(const VolParams &p)
{
    SIDE_REQUIRE(p.N >= 0 && p.B > 0 && p.C > 0 && p.H > 1 && p.W > 1, "inst_costvol: bad shape");
    SIDE_REQUIRE(p.D >= 2, "inst_costvol: D (depth candidates) must be >= 2 (reference divides by D-1)");
    SIDE_REQUIRE(p.P >= 1 && 6 * p.P <= kVolThreads, "inst_costvol: P out of range (1..%d)", kVolThreads / 6);
    SIDE_REQUIRE((long long)p.N * p.D < (1ll << 31), "inst_costvol: N*D too large");
    return SIDE_OK;
}

}  // namespace side

using namespace side;

extern "C" size_t side_inst_costvol_fast_ws_bytes(int B, int C, int H, int W, int N, int D)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || N < 0 || D <= 0) return 0;
    // gate-statistics partials only (up to 4 slots of 4 floats per 8-channel chunk); the separable kernel reads the NCHW features in place
    return sizeof(float) * (16 * (size_t)N * D * ((C + kSepCC - 1) / kSepCC) + 64);
}

extern "C" size_t side_inst_costvol_ws_bytes(int B, int C, int H, int W)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return sizeof(float) * 2 * (size_t)B * C * H * W;
}

extern "C" int side_proposal_shift(const float *left, const float *right, const float *fb, int N, int B, int D,
                                   float x_clamp, float *pro_left, float *pro_right, float *depth_bin, void *stream)
{
    SIDE_REQUIRE(N >= 0 && B > 0 && D >= 2, "side_proposal_shift: bad shape (N=%d B=%d D=%d)", N, B, D);
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(left); SIDE_REQUIRE_DEV(right); SIDE_REQUIRE_DEV(fb);
    SIDE_REQUIRE_DEV(pro_left); SIDE_REQUIRE_DEV(pro_right); SIDE_REQUIRE_DEV(depth_bin);
    proposal_shift_kernel<<<ceil_div((long long)N * D, 128), 128, 0, (cudaStream_t)stream>>>(
        left, right, fb, N, B, D, x_clamp, pro_left, pro_right, depth_bin);
    SIDE_LAUNCH_CHECK("proposal_shift_kernel");
    return SIDE_OK;
}

extern "C" int side_inst_costvol_fwd(const float *featL, const float *featR, const float *left, const float *right,
                                     const float *fb, const uint8_t *valid, float *cost, float *depth_bin,
                                     float *xcross, int N, int B, int C, int H, int W, int D, int P, float x_clamp,
                                     int flags, void *ws, size_t ws_bytes, void *stream)
{
    VolParams p{featL, featR, left, right, fb, valid, cost, depth_bin, xcross, nullptr, nullptr, nullptr,
                N, B, C, H, W, D, P, x_clamp};
    int rc = check_vol_args(p);
    if (rc) return rc;
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(featL); SIDE_REQUIRE_DEV(featR); SIDE_REQUIRE_DEV(left); SIDE_REQUIRE_DEV(right);
    SIDE_REQUIRE_DEV(fb); SIDE_REQUIRE_DEV(cost); SIDE_REQUIRE_DEV(depth_bin);
    const bool gate = flags & SIDE_VOL_GATE;
    const dim3 grid((unsigned)((long long)N * D));
    cudaStream_t st = (cudaStream_t)stream;
    if (flags & SIDE_VOL_SEPARABLE) {
        SIDE_REQUIRE(P == 16 && C % kSepCC == 0 && D <= kSepMaxD && N <= 65535,
                     "inst_costvol: the separable path needs P == 16, C %% 8 == 0, D <= %d, N <= 65535", kSepMaxD);
        SIDE_REQUIRE((long long)B * H * W * C < (1ll << 31) && W < 32768, "inst_costvol: features too large for the separable path");
        if (ws == nullptr || ws_bytes < side_inst_costvol_fast_ws_bytes(B, C, H, W, N, D) || !is_device_ptr(ws)) {
            set_error("inst_costvol: the separable path needs side_inst_costvol_fast_ws_bytes(...) bytes of device workspace");
            return SIDE_ERR_WORKSPACE;
        }
        float *partial = reinterpret_cast<float *>(ws);
        const bool want_xc = (flags & SIDE_VOL_XCROSS) && xcross != nullptr;
        static const int rows_env = [] { const char *e = getenv("SIDE_SEP_ROWS"); return e ? atoi(e) : 0; }();
        const int rows = rows_env == 16 ? 16 : 8;

#define SIDE_SEP_LAUNCH(WR, ST, AP, what)                                                                                      \
    do {                                                                                                                       \
        if (rows == 16) {                                                                                                      \
            const size_t smem = sizeof(float) * 2 * 16 * kSepRowF;                                                             \
            if ((rc = set_smem_attr((const void *)inst_costvol_sep_kernel<WR, ST, AP, 16>, smem))) return rc;                  \
            inst_costvol_sep_kernel<WR, ST, AP, 16><<<dim3((unsigned)(C / kSepCC), (unsigned)N, 1), 512, smem, st>>>(p, partial); \
        } else {                                                                                                               \
            const size_t smem = sizeof(float) * 2 * 8 * kSepRowF;                                                              \
            if ((rc = set_smem_attr((const void *)inst_costvol_sep_kernel<WR, ST, AP, 8>, smem))) return rc;                   \
            inst_costvol_sep_kernel<WR, ST, AP, 8><<<dim3((unsigned)(2 * (C / kSepCC)), (unsigned)N, 1), 256, smem, st>>>(p, partial);  \
        }                                                                                                                      \
        SIDE_LAUNCH_CHECK(what);                                                                                               \
    } while (0)
        const int nslot = (C / kSepCC) * (16 / rows);
        if (gate) {
            SIDE_SEP_LAUNCH(false, true, false, "inst_costvol_sep_kernel<stats>");
            SIDE_SEP_LAUNCH(true, false, true, "inst_costvol_sep_kernel<write, gate>");
        } else if (want_xc) {
            SIDE_SEP_LAUNCH(true, true, false, "inst_costvol_sep_kernel<write, stats>");
            sep_finish_xcross_kernel<<<ceil_div((long long)N * D, 128), 128, 0, st>>>(partial, valid, xcross, N * D, D, nslot);
            SIDE_LAUNCH_CHECK("sep_finish_xcross_kernel");
        } else {
            SIDE_SEP_LAUNCH(true, false, false, "inst_costvol_sep_kernel<write>");
        }
#undef SIDE_SEP_LAUNCH
        return SIDE_OK;
    }
    // channels-last fast path: needs the transposed copies (workspace), C % 4 == 0, P*P % 4 == 0 and the padded
    // L/R tiles in shared memory
    const size_t nhwc_smem = sizeof(AxisTap) * 6 * P + sizeof(float) * 4 * 32 + sizeof(float) * 2 * (size_t)C * (P * P + 1);
    if (ws != nullptr && ws_bytes >= side_inst_costvol_ws_bytes(B, C, H, W) && (C & 3) == 0 && ((P * P) & 3) == 0 &&
        nhwc_smem <= 200 * 1024 && (long long)B * H * W * C < (1ll << 31) && is_device_ptr(ws)) {
        float *nl = reinterpret_cast<float *>(ws), *nr = nl + (size_t)B * C * H * W;
        if ((rc = launch_nchw_to_nhwc(featL, nl, B, C, H * W, st))) return rc;
        if ((rc = launch_nchw_to_nhwc(featR, nr, B, C, H * W, st))) return rc;
#define SIDE_VOL_LAUNCH(G, F, PTV)                                                                                   \
    do {                                                                                                             \
        if ((rc = set_smem_attr((const void *)inst_costvol_fwd_nhwc_kernel<G, F, PTV>, nhwc_smem))) return rc;       \
        inst_costvol_fwd_nhwc_kernel<G, F, PTV><<<grid, kNhwcThreads, nhwc_smem, st>>>(p, nl, nr);                    \
    } while (0)
        const bool fast = flags & SIDE_VOL_FMA;
        if (P == 16) {
            if (gate && fast) SIDE_VOL_LAUNCH(true, true, 16);
            else if (gate) SIDE_VOL_LAUNCH(true, false, 16);
            else if (fast) SIDE_VOL_LAUNCH(false, true, 16);
            else SIDE_VOL_LAUNCH(false, false, 16);
        } else {
            if (gate && fast) SIDE_VOL_LAUNCH(true, true, 0);
            else if (gate) SIDE_VOL_LAUNCH(true, false, 0);
            else if (fast) SIDE_VOL_LAUNCH(false, true, 0);
            else SIDE_VOL_LAUNCH(false, false, 0);
        }
#undef SIDE_VOL_LAUNCH
        SIDE_LAUNCH_CHECK("inst_costvol_fwd_nhwc_kernel");
        return SIDE_OK;
    }
    const bool stage = gate && vol_smem_bytes(C, P, true) <= 200 * 1024;
    const size_t smem = vol_smem_bytes(C, P, stage);
    if (gate && stage) {
        if ((rc = set_smem_attr((const void *)inst_costvol_fwd_kernel<true, true>, smem))) return rc;
        inst_costvol_fwd_kernel<true, true><<<grid, kVolThreads, smem, st>>>(p);
    } else if (gate) {
        inst_costvol_fwd_kernel<true, false><<<grid, kVolThreads, smem, st>>>(p);
    } else {
        inst_costvol_fwd_kernel<false, false><<<grid, kVolThreads, smem, st>>>(p);
    }
    SIDE_LAUNCH_CHECK("inst_costvol_fwd_kernel");
    return SIDE_OK;
}

extern "C" int side_inst_costvol_bwd(const float *featL, const float *featR, const float *left, const float *right,
                                     const float *fb, const uint8_t *valid, const float *gcost, float *gfeatL,
                                     float *gfeatR, int N, int B, int C, int H, int W, int D, int P, float x_clamp,
                                     int flags, void *stream)
{
    VolParams p{featL, featR, left, right, fb, valid, nullptr, nullptr, nullptr, gcost, gfeatL, gfeatR,
                N, B, C, H, W, D, P, x_clamp};
    int rc = check_vol_args(p);
    if (rc) return rc;
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(featL); SIDE_REQUIRE_DEV(featR); SIDE_REQUIRE_DEV(left); SIDE_REQUIRE_DEV(right);
    SIDE_REQUIRE_DEV(fb); SIDE_REQUIRE_DEV(gcost); SIDE_REQUIRE_DEV(gfeatL); SIDE_REQUIRE_DEV(gfeatR);
    const bool gate = flags & SIDE_VOL_GATE;
    cudaStream_t st0 = (cudaStream_t)stream;
    // ungated volume on the fused geometry: register-accumulating gather kernel (no atomics in the x pass)
    if (!gate && !(flags & SIDE_VOL_BWD_SCALAR) && P == 16 && C % kBgCC == 0 && D <= kSepMaxD && H <= kBgMaxRows && N <= 65535) {
        const size_t smem = sizeof(float) * (size_t)kBgStages * kBgStageF;
        if ((rc = set_smem_attr((const void *)inst_costvol_bwd_gather_kernel, smem))) return rc;
        inst_costvol_bwd_gather_kernel<<<dim3((unsigned)(C / kBgCC), (unsigned)N), kBgThreads, smem, st0>>>(p);
        SIDE_LAUNCH_CHECK("inst_costvol_bwd_gather_kernel");
        return SIDE_OK;
    }
    const bool stage = gate && vol_smem_bytes(C, P, true) <= 200 * 1024;
    const size_t smem = vol_smem_bytes(C, P, stage);
    const dim3 grid((unsigned)((long long)N * D));
    cudaStream_t st = (cudaStream_t)stream;
    if (gate && stage) {
        if ((rc = set_smem_attr((const void *)inst_costvol_bwd_kernel<true, true>, smem))) return rc;
        inst_costvol_bwd_kernel<true, true><<<grid, kVolThreads, smem, st>>>(p);
    } else if (gate) {
        inst_costvol_bwd_kernel<true, false><<<grid, kVolThreads, smem, st>>>(p);
    } else {
        inst_costvol_bwd_kernel<false, false><<<grid, kVolThreads, smem, st>>>(p);
    }
    SIDE_LAUNCH_CHECK("inst_costvol_bwd_kernel");
    return SIDE_OK;
}

extern "C" int side_xcross_gate_fwd(const float *cost, float *out, float *xcross, int N, int C, int D, int P,
                                    void *stream)
{
    SIDE_REQUIRE(N >= 0 && C > 0 && D > 0 && P > 0, "side_xcross_gate_fwd: bad shape");
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(cost); SIDE_REQUIRE_DEV(out);
    xcross_gate_kernel<false><<<(unsigned)((long long)N * D), 256, 0, (cudaStream_t)stream>>>(cost, nullptr, out,
                                                                                            xcross, C, D, P * P);
    SIDE_LAUNCH_CHECK("xcross_gate_kernel<fwd>");
    return SIDE_OK;
}

extern "C" int side_xcross_gate_bwd(const float *cost, const float *gout, float *gcost, int N, int C, int D, int P,
                                    void *stream)
{
    SIDE_REQUIRE(N >= 0 && C > 0 && D > 0 && P > 0, "side_xcross_gate_bwd: bad shape");
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(cost); SIDE_REQUIRE_DEV(gout); SIDE_REQUIRE_DEV(gcost);
    xcross_gate_kernel<true><<<(unsigned)((long long)N * D), 256, 0, (cudaStream_t)stream>>>(cost, gout, gcost,
                                                                                           nullptr, C, D, P * P);
    SIDE_LAUNCH_CHECK("xcross_gate_kernel<bwd>");
    return SIDE_OK;
}

// ---- backward with workspace: the gated volume's backward as  recompute raw volume (separable forward) -> gate backward
//      in place -> separable gather backward; without the gate it is the gather kernel alone ----
extern "C" size_t side_inst_costvol_bwd_fast_ws_bytes(int B, int C, int H, int W, int N, int D, int flags)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || N <= 0 || D <= 0 || !(flags & SIDE_VOL_GATE)) return 0;
    const size_t fwd = (side_inst_costvol_fast_ws_bytes(B, C, H, W, N, D) + 255) / 256 * 256;
    return fwd + sizeof(float) * ((size_t)N * 3 * C * D * 256 + (size_t)N * D);
}

extern "C" int side_inst_costvol_bwd_fast(const float *featL, const float *featR, const float *left, const float *right,
                                          const float *fb, const uint8_t *valid, const float *gcost, float *gfeatL,
                                          float *gfeatR, int N, int B, int C, int H, int W, int D, int P, float x_clamp,
                                          int flags, void *ws, size_t ws_bytes, void *stream)
{
    const bool fast_ok = P == 16 && C % kBgCC == 0 && C % kSepCC == 0 && D <= kSepMaxD && H <= kBgMaxRows && N <= 65535 && N > 0 &&
                         !(flags & SIDE_VOL_BWD_SCALAR);
    if (!(flags & SIDE_VOL_GATE) || !fast_ok)
        return side_inst_costvol_bwd(featL, featR, left, right, fb, valid, gcost, gfeatL, gfeatR, N, B, C, H, W, D, P, x_clamp,
                                     flags, stream);
    const size_t need = side_inst_costvol_bwd_fast_ws_bytes(B, C, H, W, N, D, flags);
    if (ws == nullptr || ws_bytes < need || !is_device_ptr(ws)) {
        set_error("side_inst_costvol_bwd_fast: needs side_inst_costvol_bwd_fast_ws_bytes(...) = %zu bytes of device workspace", need);
        return SIDE_ERR_WORKSPACE;
    }
    SIDE_REQUIRE_DEV(gcost);
    const size_t fwd = (side_inst_costvol_fast_ws_bytes(B, C, H, W, N, D) + 255) / 256 * 256;
    float *raw = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(ws) + fwd);
    float *dbin = raw + (size_t)N * 3 * C * D * 256;
    int rc = side_inst_costvol_fwd(featL, featR, left, right, fb, valid, raw, dbin, nullptr, N, B, C, H, W, D, P, x_clamp,
                                   SIDE_VOL_SEPARABLE, ws, fwd, stream);
    if (rc) return rc;
    // gradient w.r.t. the raw volume, in place over the recomputed volume (each element is read before it is overwritten)
    xcross_gate_kernel<true><<<(unsigned)((long long)N * D), 256, 0, (cudaStream_t)stream>>>(raw, gcost, raw, nullptr, C, D, 256);
    SIDE_LAUNCH_CHECK("xcross_gate_kernel<bwd>");
    return side_inst_costvol_bwd(featL, featR, left, right, fb, valid, raw, gfeatL, gfeatR, N, B, C, H, W, D, P, x_clamp,
                                 flags & ~SIDE_VOL_GATE, stream);
}
