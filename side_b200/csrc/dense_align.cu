// dense_align.cu -- dense photometric alignment of the post-process path.
//
// Reference: src/lib/dense_align/dense_align.py.  align_parallel (:240-312) normalises the two raw images on the host
// with numpy, up-samples them 2x (F.interpolate, bilinear), builds per-RoI pixel samples with a Python loop over
// RoIs (sample, :14-70, Box3d ray/box intersection box_3d.py:9-102), and scores 50 + 20 depth hypotheses per RoI by
// 2 x (iter x rois x pixels) F.grid_sample calls over tensors expanded to the full (iter x rois x pixels) size
// (enumeration_depth, :175-237).  Here:
//   da_prep_u8_kernel   uint8 HWC image -> ((x / 255) - mean) / std -> 2x bilinear -> PACKED [2H][2W] float4 texels
//                       (the three channels of one sample position are one 16-byte load).
//   da_sample_kernel    one CTA per RoI: box planes, nearest vertex, ray/plane hits, inside-box mask in the reference's
//                       float32 operation order; valid pixels are compacted in row-major order (ballot + prefix), so
//                       the result is the reference's (rois x pixels x 3, rois x pixels) pair with a fixed capacity.
//   da_enum_kernel      gather-and-reduce: CTA = (RoI, 10 hypotheses).  The left sample and the vertical interpolation
//                       are computed once per pixel; the right image is sampled per hypothesis; 10 running sums per
//                       thread, shuffle + shared reduction in a fixed order (deterministic).  No expanded grids.
//   da_argmin_kernel    first minimum over the hypotheses (torch.min semantics) -> best depth per RoI.
#include <algorithm>
#include "common.cuh"

namespace side {

constexpr int kDaThreads = 256;
constexpr int kDaIters = 10;        // hypotheses per CTA of the enumeration kernel

// ------------------------------------------------------------------------------------------------------------------
// image preparation
// ------------------------------------------------------------------------------------------------------------------
// upsample_bilinear2d with scale_factor = 2, align_corners = False (ATen area_pixel_compute_source_index):
// src = 0.5 * (dst + 0.5) - 0.5, clamped at 0; lambda1 = src - floor(src)
__device__ __forceinline__ void up2_src(int dst, int size, int &i0, int &ip, float &l0, float &l1)
{
    float src = __fsub_rn(__fmul_rn(0.5f, __fadd_rn((float)dst, 0.5f)), 0.5f);
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    ip = (i0 < size - 1) ? 1 : 0;
    l1 = __fsub_rn(src, (float)i0);
    l0 = __fsub_rn(1.f, l1);
}

template <bool FROM_U8>
__global__ void da_prep_kernel(const unsigned char *__restrict__ u8, const float *__restrict__ planar, float4 *__restrict__ out,
                               int H, int W, float m0, float m1, float m2, float s0, float s1, float s2)
{
    const int H2 = 2 * H, W2 = 2 * W;
    const long long total = (long long)H2 * W2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / W2), x = (int)(i - (long long)y * W2);
        int y0, yp, x0, xp;
        float hy0, hy1, wx0, wx1;
        up2_src(y, H, y0, yp, hy0, hy1);
        up2_src(x, W, x0, xp, wx0, wx1);
        const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float p00, p01, p10, p11;
            if (FROM_U8) {
                auto px = [&](int yy, int xx) {
                    const float v = __fdiv_rn((float)u8[((size_t)yy * W + xx) * 3 + c], 255.f);       // img.astype(f32) / 255.
                    return __fdiv_rn(__fsub_rn(v, mean[c]), sd[c]);                                  // (im - mean) / std
                };
                p00 = px(y0, x0); p01 = px(y0, x0 + xp); p10 = px(y0 + yp, x0); p11 = px(y0 + yp, x0 + xp);
            } else {
                const float *pc = planar + (size_t)c * H * W;
                p00 = __ldg(pc + (size_t)y0 * W + x0); p01 = __ldg(pc + (size_t)y0 * W + x0 + xp);
                p10 = __ldg(pc + (size_t)(y0 + yp) * W + x0); p11 = __ldg(pc + (size_t)(y0 + yp) * W + x0 + xp);
            }
            // h0 * (w0 * p00 + w1 * p01) + h1 * (w0 * p10 + w1 * p11), separately rounded (UpSampleBilinear2d CPU order)
            const float top = __fadd_rn(__fmul_rn(wx0, p00), __fmul_rn(wx1, p01));
            const float bot = __fadd_rn(__fmul_rn(wx0, p10), __fmul_rn(wx1, p11));
            o[c] = __fadd_rn(__fmul_rn(hy0, top), __fmul_rn(hy1, bot));
        }
        out[i] = make_float4(o[0], o[1], o[2], 0.f);
    }
}

// planar [3][H][W] -> packed [H][W] float4 (for the drop-in enumeration_depth, whose images arrive as NCHW tensors)
__global__ void da_pack_kernel(const float *__restrict__ planar, float4 *__restrict__ out, long long HW)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x)
        out[i] = make_float4(__ldg(planar + i), __ldg(planar + HW + i), __ldg(planar + 2 * HW + i), 0.f);
}
__global__ void da_unpack_kernel(const float4 *__restrict__ packed, float *__restrict__ planar, long long HW)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = packed[i];
        planar[i] = v.x; planar[HW + i] = v.y; planar[2 * HW + i] = v.z;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// sample(): per-RoI pixel set with the depth offset of the visible box surface (dense_align.py:14-70, box_3d.py)
// ------------------------------------------------------------------------------------------------------------------
struct DaBox {
    float T[3], R[9];          // T_c_o, R_c_o (row major)
    float plane[3][4];         // the three visible planes of the nearest vertex, in test order
    float lo[3], hi[3];        // object-frame bounds with the reference's 0.01 slack
    int r0, r1, rs, c0, c1, cs;  // python slice start / stop / step of rows and columns (already normalised to the image)
    int nrows, ncols;
};

// Python's slice.indices() for a positive step
__device__ inline void py_slice(int start, int stop, int len, int &s, int &e)
{
    if (start < 0) { start += len; if (start < 0) start = 0; } else if (start > len) start = len;
    if (stop < 0) { stop += len; if (stop < 0) stop = 0; } else if (stop > len) stop = len;
    s = start; e = stop;
}

__device__ inline void da_make_plane(const float *p1, const float *p2, const float *p3, float *pl)
{
    // creatPlane (box_3d.py:32-41): normal = cross(p2 - p1, p3 - p1); d = -n0*p1[0] - n1*p1[1] - n2*p1[2]
    const float a0 = __fsub_rn(p2[0], p1[0]), a1 = __fsub_rn(p2[1], p1[1]), a2 = __fsub_rn(p2[2], p1[2]);
    const float b0 = __fsub_rn(p3[0], p1[0]), b1 = __fsub_rn(p3[1], p1[1]), b2 = __fsub_rn(p3[2], p1[2]);
    const float n0 = __fsub_rn(__fmul_rn(a1, b2), __fmul_rn(a2, b1));
    const float n1 = __fsub_rn(__fmul_rn(a2, b0), __fmul_rn(a0, b2));
    const float n2 = __fsub_rn(__fmul_rn(a0, b1), __fmul_rn(a1, b0));
    pl[0] = n0; pl[1] = n1; pl[2] = n2;
    pl[3] = __fsub_rn(__fsub_rn(__fmul_rn(-n0, p1[0]), __fmul_rn(n1, p1[1])), __fmul_rn(n2, p1[2]));
}

__global__ void __launch_bounds__(kDaThreads) da_sample_kernel(const float *__restrict__ box, const float *__restrict__ borders,
                                                              const float *__restrict__ poses, float f, float cx, float cy,
                                                              int f_h, int f_w, int cap, float *__restrict__ uvz,
                                                              float *__restrict__ weight, int *__restrict__ count)
{
    __shared__ DaBox sb;
    __shared__ int s_warp[kDaThreads / 32];
    __shared__ int s_base;
    const int roi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const float *bx = box + 4 * roi, *bd = borders + 2 * roi, *ps = poses + 7 * roi;
        // slice bounds (dense_align.py:42-45); every operation is a float32 tensor op followed by int() truncation
        const int width = max((int)__fdiv_rn(__fsub_rn(bd[1], bd[0]), 56.f), 1);
        const int height = max((int)__fdiv_rn(__fsub_rn(bx[3], bx[1]), 56.f), 1);
        const int rs = (int)__fadd_rn(__fdiv_rn(__fadd_rn(bx[1], bx[3]), 2.f), 0.5f);
        const int re = (int)__fadd_rn(__fsub_rn(bx[3], __fmul_rn(__fsub_rn(bx[3], bx[1]), 0.1f)), 0.5f);
        const int cs = (int)__fadd_rn(bd[0], 0.5f), ce = (int)__fadd_rn(bd[1], 0.5f);
        py_slice(rs, re, f_h, sb.r0, sb.r1);
        py_slice(cs, ce, f_w, sb.c0, sb.c1);
        sb.rs = height; sb.cs = width;
        sb.nrows = sb.r1 > sb.r0 ? (sb.r1 - sb.r0 + height - 1) / height : 0;
        sb.ncols = sb.c1 > sb.c0 ? (sb.c1 - sb.c0 + width - 1) / width : 0;
        // Box3d.__init__ (box_3d.py:10-56)
        const float cs_ = (float)cos((double)ps[6]), sn_ = (float)sin((double)ps[6]);
        const float R[9] = {cs_, 0.f, sn_, 0.f, 1.f, 0.f, -sn_, 0.f, cs_};
        for (int i = 0; i < 9; ++i) sb.R[i] = R[i];
        for (int i = 0; i < 3; ++i) sb.T[i] = ps[i];
        const float hw = __fdiv_rn(ps[3], 2.f), hl = __fdiv_rn(ps[5], 2.f), hh = ps[4];
        const float Po[8][3] = {{-hw, 0.f, -hl}, {-hw, 0.f, hl}, {hw, 0.f, hl}, {hw, 0.f, -hl},
                                {-hw, -hh, -hl}, {-hw, -hh, hl}, {hw, -hh, hl}, {hw, -hh, -hl}};
        float Pc[8][3];
        int nearest = 0;
        float nd = 100000000.f;
        for (int i = 0; i < 8; ++i) {
            for (int r = 0; r < 3; ++r)        // torch.mm(R, P_o[i]) + T: three products summed left to right
                Pc[i][r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3 * r], Po[i][0]), __fmul_rn(R[3 * r + 1], Po[i][1])),
                                               __fmul_rn(R[3 * r + 2], Po[i][2])), ps[r]);
            const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(Pc[i][0], Pc[i][0]), __fmul_rn(Pc[i][1], Pc[i][1])),
                                                   __fmul_rn(Pc[i][2], Pc[i][2])));
            if (nrm < nd) { nd = nrm; nearest = i; }
        }
        float planes[6][4];
        da_make_plane(Pc[0], Pc[3], Pc[4], planes[0]);   // front
        da_make_plane(Pc[2], Pc[3], Pc[6], planes[1]);   // right
        da_make_plane(Pc[1], Pc[2], Pc[5], planes[2]);   // back
        da_make_plane(Pc[0], Pc[1], Pc[4], planes[3]);   // left
        da_make_plane(Pc[0], Pc[1], Pc[2], planes[4]);   // bottom
        da_make_plane(Pc[4], Pc[5], Pc[6], planes[5]);   // top
        const int group[8][3] = {{0, 3, 4}, {2, 3, 4}, {1, 2, 4}, {0, 1, 4}, {0, 3, 5}, {2, 3, 5}, {1, 2, 5}, {0, 1, 5}};
        for (int k = 0; k < 3; ++k)
            for (int j = 0; j < 4; ++j) sb.plane[k][j] = planes[group[nearest][k]][j];
        // mask_out_box bounds (box_3d.py:69-74): P_o[4] - eps .. P_o[2] + eps
        sb.lo[0] = __fsub_rn(-hw, 0.01f); sb.lo[1] = __fsub_rn(-hh, 0.01f); sb.lo[2] = __fsub_rn(-hl, 0.01f);
        sb.hi[0] = __fadd_rn(hw, 0.01f);  sb.hi[1] = __fadd_rn(0.f, 0.01f); sb.hi[2] = __fadd_rn(hl, 0.01f);
        s_base = 0;
    }
    __syncthreads();
    const int npix = sb.nrows * sb.ncols;
    float *uo = uvz + (size_t)roi * cap * 3, *wo = weight + (size_t)roi * cap;
    for (int base = 0; base < npix; base += kDaThreads) {
        const int i = base + tid;
        bool valid = false;
        float u = 0.f, v = 0.f, z = 0.f;
        if (i < npix) {
            const int r = i / sb.ncols, c = i - r * sb.ncols;
            u = (float)(sb.c0 + c * sb.cs);
            v = (float)(sb.r0 + r * sb.rs);
            const float nu = __fdiv_rn(__fsub_rn(u, cx), f), nv = __fdiv_rn(__fsub_rn(v, cy), f);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (valid) break;                              // mask_out_box only touches pixels whose flag is still 0
                const float *pl = sb.plane[k];
                float t = __fadd_rn(__fadd_rn(__fmul_rn(nu, pl[0]), __fmul_rn(nv, pl[1])), __fmul_rn(1.f, pl[2]));
                t = __fmul_rn(-__frcp_rn(t), pl[3]);
                const float ic0 = __fsub_rn(__fmul_rn(nu, t), sb.T[0]), ic1 = __fsub_rn(__fmul_rn(nv, t), sb.T[1]),
                            ic2 = __fsub_rn(__fmul_rn(1.f, t), sb.T[2]);
                // object frame: R^T * ic
                const float o0 = __fadd_rn(__fadd_rn(__fmul_rn(sb.R[0], ic0), __fmul_rn(sb.R[3], ic1)), __fmul_rn(sb.R[6], ic2));
                const float o1 = __fadd_rn(__fadd_rn(__fmul_rn(sb.R[1], ic0), __fmul_rn(sb.R[4], ic1)), __fmul_rn(sb.R[7], ic2));
                const float o2 = __fadd_rn(__fadd_rn(__fmul_rn(sb.R[2], ic0), __fmul_rn(sb.R[5], ic1)), __fmul_rn(sb.R[8], ic2));
                z = ic2;
                valid = o0 >= sb.lo[0] && o1 >= sb.lo[1] && o2 >= sb.lo[2] && o0 <= sb.hi[0] && o1 <= sb.hi[1] && o2 <= sb.hi[2];
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        off += __popc(bal & ((1u << lane) - 1u));
        if (valid && off < cap) {
            uo[3 * off] = u; uo[3 * off + 1] = v; uo[3 * off + 2] = z;
            wo[off] = 1.f;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < kDaThreads / 32; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    const int n = s_base;
    if (tid == 0) count[roi] = n;                              // may exceed cap: the caller checks
    for (int i = min(n, cap) + tid; i < cap; i += kDaThreads) {
        uo[3 * i] = 0.f; uo[3 * i + 1] = 0.f; uo[3 * i + 2] = 0.f;
        wo[i] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// enumeration_depth (dense_align.py:175-237)
// ------------------------------------------------------------------------------------------------------------------
// F.grid_sample coordinate: normalised g -> pixel position, padding_mode='border' (ATen GridSampler.h)
template <bool ALIGN>
__device__ __forceinline__ float da_unnorm(float g, int size)
{
    float x = ALIGN ? __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), (float)(size - 1))
                    : __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), 1.f), 2.f);
    return fminf((float)(size - 1), fmaxf(x, 0.f));
}

template <bool ALIGN>
__global__ void __launch_bounds__(kDaThreads) da_enum_kernel(const float4 *__restrict__ imL, const float4 *__restrict__ imR,
                                                            const float *__restrict__ uvz, const float *__restrict__ weight,
                                                            const float *__restrict__ depth_enum, float fb, int rois,
                                                            int pixels, int iters, int H, int W, float half_w, float half_h,
                                                            float *__restrict__ err_sum)
{
    __shared__ float s_red[kDaThreads / 32][kDaIters];
    const int roi = blockIdx.x, it0 = blockIdx.y * kDaIters, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nit = min(kDaIters, iters - it0);
    float inv_dis[kDaIters], acc[kDaIters];
#pragma unroll
    for (int k = 0; k < kDaIters; ++k) {
        acc[k] = 0.f;
        // dis = depth.reciprocal() * fb; the per-pixel term needs dis.reciprocal()
        const float d = k < nit ? __ldg(depth_enum + (size_t)(it0 + k) * rois + roi) : 1.f;
        inv_dis[k] = __frcp_rn(__fmul_rn(__frcp_rn(d), fb));
    }
    const float *up = uvz + (size_t)roi * pixels * 3, *wp = weight + (size_t)roi * pixels;
    for (int p = tid; p < pixels; p += kDaThreads) {
        const float wgt = __ldg(wp + p);
        if (wgt == 0.f) continue;
        const float u = __ldg(up + 3 * p), v = __ldg(up + 3 * p + 1), z = __ldg(up + 3 * p + 2);
        // vertical interpolation is shared by the left sample and every right sample
        const float gy = __fdiv_rn(__fsub_rn(v, half_h), half_h);
        const float y = da_unnorm<ALIGN>(gy, H);
        const float yn = floorf(y);
        const float wn = __fsub_rn(y, yn), ws = __fsub_rn(1.f, wn);      // weight of the lower / upper row
        const int iy0 = (int)yn, iy1 = min(iy0 + 1, H - 1);
        const bool y1_in = iy0 + 1 <= H - 1;
        const float4 *rowL0 = imL + (size_t)iy0 * W, *rowL1 = imL + (size_t)iy1 * W;
        const float4 *rowR0 = imR + (size_t)iy0 * W, *rowR1 = imR + (size_t)iy1 * W;
        auto sample = [&](const float4 *r0, const float4 *r1, float g, float &c0, float &c1, float &c2) {
            const float x = da_unnorm<ALIGN>(g, W);
            const float xw = floorf(x);
            const float we = __fsub_rn(x, xw), ww = __fsub_rn(1.f, we);
            const int ix0 = (int)xw, ix1 = min(ix0 + 1, W - 1);
            const bool x1_in = ix0 + 1 <= W - 1;
            const float nw = __fmul_rn(ws, ww), ne = x1_in ? __fmul_rn(ws, we) : 0.f;
            const float sw = y1_in ? __fmul_rn(wn, ww) : 0.f, se = (x1_in && y1_in) ? __fmul_rn(wn, we) : 0.f;
            const float4 a = __ldg(r0 + ix0), b = __ldg(r0 + ix1), c = __ldg(r1 + ix0), d = __ldg(r1 + ix1);
            c0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.x, nw), __fmul_rn(b.x, ne)), __fmul_rn(c.x, sw)), __fmul_rn(d.x, se));
            c1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.y, nw), __fmul_rn(b.y, ne)), __fmul_rn(c.y, sw)), __fmul_rn(d.y, se));
            c2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.z, nw), __fmul_rn(b.z, ne)), __fmul_rn(c.z, sw)), __fmul_rn(d.z, se));
        };
        float l0, l1, l2;
        sample(rowL0, rowL1, __fdiv_rn(__fsub_rn(u, half_w), half_w), l0, l1, l2);
        const float zf = __fdiv_rn(z, fb);
#pragma unroll
        for (int k = 0; k < kDaIters; ++k) {
            if (k < nit) {
                // global_delta_d = (z / fb + dis.reciprocal()).reciprocal(); u_right = u - global_delta_d
                const float dd = __frcp_rn(__fadd_rn(zf, inv_dis[k]));
                const float g = __fdiv_rn(__fsub_rn(__fsub_rn(u, dd), half_w), half_w);
                float r0, r1, r2;
                sample(rowR0, rowR1, g, r0, r1, r2);
                acc[k] += (fabsf(__fsub_rn(l0, r0)) + fabsf(__fsub_rn(l1, r1)) + fabsf(__fsub_rn(l2, r2))) * wgt;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kDaIters; ++k) {
        const float s = warp_sum(acc[k]);
        if (lane == 0) s_red[warp][k] = s;
    }
    __syncthreads();
    if (tid < nit) {
        float s = 0.f;
        for (int w = 0; w < kDaThreads / 32; ++w) s += s_red[w][tid];
        err_sum[(size_t)(it0 + tid) * rois + roi] = s;
    }
}

// torch.min(error_sum, 0): first minimum; best_depth = depth_enum[idx, roi]
__global__ void da_argmin_kernel(const float *__restrict__ err_sum, const float *__restrict__ depth_enum, int rois, int iters,
                                 float *__restrict__ best_depth, int *__restrict__ best_idx)
{
    const int roi = blockIdx.x * blockDim.x + threadIdx.x;
    if (roi >= rois) return;
    float best = err_sum[roi];
    int bi = 0;
    for (int i = 1; i < iters; ++i) {
        const float e = err_sum[(size_t)i * rois + roi];
        if (e < best || (e != e && best == best)) { best = e; bi = i; }     // NaN propagates like torch.min
    }
    best_depth[roi] = depth_enum[(size_t)bi * rois + roi];
    if (best_idx) best_idx[roi] = bi;
}

}  // namespace side

using namespace side;

extern "C" int side_dense_align_prep_u8(const unsigned char *img_hwc, float *out_packed, int H, int W, const float *mean3,
                                        const float *std3, void *stream)
{
    SIDE_REQUIRE_DEV(img_hwc);
    SIDE_REQUIRE_DEV(out_packed);
    SIDE_REQUIRE(H > 0 && W > 0 && mean3 && std3, "side_dense_align_prep_u8: bad arguments");
    const long long total = 4ll * H * W;
    da_prep_kernel<true><<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        img_hwc, nullptr, reinterpret_cast<float4 *>(out_packed), H, W, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2]);
    SIDE_LAUNCH_CHECK("da_prep_kernel<u8>");
    return SIDE_OK;
}

extern "C" int side_dense_align_up2_pack(const float *img_chw, float *out_packed, int H, int W, void *stream)
{
    SIDE_REQUIRE_DEV(img_chw);
    SIDE_REQUIRE_DEV(out_packed);
    SIDE_REQUIRE(H > 0 && W > 0, "side_dense_align_up2_pack: bad arguments");
    const long long total = 4ll * H * W;
    da_prep_kernel<false><<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        nullptr, img_chw, reinterpret_cast<float4 *>(out_packed), H, W, 0.f, 0.f, 0.f, 1.f, 1.f, 1.f);
    SIDE_LAUNCH_CHECK("da_prep_kernel<planar>");
    return SIDE_OK;
}

extern "C" int side_dense_align_pack(const float *img_chw, float *out_packed, int H, int W, void *stream)
{
    SIDE_REQUIRE_DEV(img_chw);
    SIDE_REQUIRE_DEV(out_packed);
    const long long HW = (long long)H * W;
    da_pack_kernel<<<(unsigned)std::min<long long>((HW + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        img_chw, reinterpret_cast<float4 *>(out_packed), HW);
    SIDE_LAUNCH_CHECK("da_pack_kernel");
    return SIDE_OK;
}

extern "C" int side_dense_align_unpack(const float *packed, float *img_chw, int H, int W, void *stream)
{
    SIDE_REQUIRE_DEV(img_chw);
    SIDE_REQUIRE_DEV(packed);
    const long long HW = (long long)H * W;
    da_unpack_kernel<<<(unsigned)std::min<long long>((HW + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(packed), img_chw, HW);
    SIDE_LAUNCH_CHECK("da_unpack_kernel");
    return SIDE_OK;
}

extern "C" int side_dense_align_sample(const float *box_left, const float *borders, const float *poses, int rois, float f, float cx,
                                       float cy, int f_h, int f_w, int cap, float *uvz, float *weight, int *count, void *stream)
{
    if (rois == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(box_left);
    SIDE_REQUIRE_DEV(borders);
    SIDE_REQUIRE_DEV(poses);
    SIDE_REQUIRE_DEV(uvz);
    SIDE_REQUIRE_DEV(weight);
    SIDE_REQUIRE_DEV(count);
    SIDE_REQUIRE(rois > 0 && cap > 0 && f_h > 0 && f_w > 0, "side_dense_align_sample: bad sizes");
    da_sample_kernel<<<rois, kDaThreads, 0, (cudaStream_t)stream>>>(box_left, borders, poses, f, cx, cy, f_h, f_w, cap, uvz, weight,
                                                                  count);
    SIDE_LAUNCH_CHECK("da_sample_kernel");
    return SIDE_OK;
}

extern "C" int side_dense_align_enum(const float *imL_packed, const float *imR_packed, const float *uvz, const float *weight,
                                     const float *depth_enum, float fb, int rois, int pixels, int iters, int H, int W, int flags,
                                     float *err_sum, float *best_depth, int *best_idx, void *stream)
{
    if (rois == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(imL_packed);
    SIDE_REQUIRE_DEV(imR_packed);
    SIDE_REQUIRE_DEV(depth_enum);
    SIDE_REQUIRE_DEV(err_sum);
    SIDE_REQUIRE_DEV(best_depth);
    SIDE_REQUIRE(rois > 0 && iters > 0 && pixels >= 0 && H > 0 && W > 0, "side_dense_align_enum: bad sizes");
    if (pixels > 0) {
        SIDE_REQUIRE_DEV(uvz);
        SIDE_REQUIRE_DEV(weight);
    }
    const cudaStream_t st = (cudaStream_t)stream;
    // grid = (u - f_w / 2) / (f_w / 2) with f_w = W - 1 (dense_align.py:194-201): python doubles cast to float at the op
    const float half_w = (float)(((double)W - 1.0) / 2.0), half_h = (float)(((double)H - 1.0) / 2.0);
    const dim3 grid((unsigned)rois, (unsigned)ceil_div(iters, kDaIters));
    SIDE_REQUIRE(grid.y <= 65535u, "side_dense_align_enum: too many hypotheses");
    if (flags & SIDE_DA_ALIGN_CORNERS)
        da_enum_kernel<true><<<grid, kDaThreads, 0, st>>>(reinterpret_cast<const float4 *>(imL_packed),
                                                          reinterpret_cast<const float4 *>(imR_packed), uvz, weight, depth_enum, fb,
                                                          rois, pixels, iters, H, W, half_w, half_h, err_sum);
    else
        da_enum_kernel<false><<<grid, kDaThreads, 0, st>>>(reinterpret_cast<const float4 *>(imL_packed),
                                                           reinterpret_cast<const float4 *>(imR_packed), uvz, weight, depth_enum, fb,
                                                           rois, pixels, iters, H, W, half_w, half_h, err_sum);
    SIDE_LAUNCH_CHECK("da_enum_kernel");
    da_argmin_kernel<<<ceil_div(rois, 128), 128, 0, st>>>(err_sum, depth_enum, rois, iters, best_depth, best_idx);
    SIDE_LAUNCH_CHECK("da_argmin_kernel");
    return SIDE_OK;
}
