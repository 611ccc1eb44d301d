// upsample.cu -- depth-wise ConvTranspose2d of the DLA up-sampling neck.
//
// Reference: IDAUp.up_k = nn.ConvTranspose2d(o, o, 2f, stride=f, padding=f//2, groups=o, bias=False) initialised to
// bilinear interpolation (feature_extraction_dla34.py:333-342, 370-373), executed between every proj/node DCN pair.
// cuDNN runs it through a generic grouped direct-convolution kernel that measured ~9 ms per call on B200 at
// micro-batch 4 (41 % of the whole inference step, profiles/r1_launches_step_fp32.md) although the layer moves only
// tens of MB.  Here: one thread per output pixel, at most ceil(k/s)^2 multiply-adds, fully coalesced along W;
// HBM-bound (reads x once, writes y once).  Backward: gather form for grad_input, block reduction for grad_weight.
#include "tc_common.cuh"

namespace side {

__global__ void __launch_bounds__(256) dw_deconv_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                           float *__restrict__ y, int C, int H, int W, int Ho, int Wo, int k,
                                                           int s, int p)
{
    const int bc = blockIdx.z, c = bc % C;
    const int oy = blockIdx.y;
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    if (ox >= Wo) return;
    const float *xp = x + (size_t)bc * H * W;
    const float *wp = w + (size_t)c * k * k;
    float acc = 0.f;
    // taps with (oy + p - ky) % s == 0, visited in increasing ky (= decreasing iy)
    for (int ky = (oy + p) % s; ky < k; ky += s) {
        const int iy = (oy + p - ky) / s;
        if (iy < 0 || iy >= H) continue;
        for (int kx = (ox + p) % s; kx < k; kx += s) {
            const int ix = (ox + p - kx) / s;
            if (ix < 0 || ix >= W) continue;
            acc = fmaf(__ldg(xp + iy * W + ix), __ldg(wp + ky * k + kx), acc);
        }
    }
    y[((size_t)bc * Ho + oy) * Wo + ox] = acc;
}

// Vectorised forward for Wo % 4 == 0: a thread produces 4 consecutive outputs of one row (one 16-byte store), a CTA of 256
// threads covers 1024 outputs of the flattened (row, x) plane of one (b, c) image -- the per-row launch above spends its
// time scheduling ~50 000 CTAs of 80-320 outputs each.  Same tap order (increasing ky, kx), so the results are identical.
template <int ST>     // ST > 0: compile-time stride with k = 2 * ST, p = ST / 2 (the IDAUp configuration): shifts, no divisions
__global__ void __launch_bounds__(256) dw_deconv_fwd_v4_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                              float *__restrict__ y, int C, int H, int W, int Ho, int Wo, int k_rt,
                                                              int s_rt, int p_rt)
{
    const int s = ST > 0 ? ST : s_rt, k = ST > 0 ? 2 * ST : k_rt, p = ST > 0 ? ST / 2 : p_rt;
    const int bc = blockIdx.y, c = bc % C;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;       // index of the 4-output group in the (Ho, Wo/4) plane
    const int wq = Wo >> 2;
    if (q >= Ho * wq) return;
    const int oy = q / wq, ox0 = (q - oy * wq) << 2;
    const float *xp = x + (size_t)bc * H * W;
    const float *wp = w + (size_t)c * k * k;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ky = (oy + p) % s; ky < k; ky += s) {
        const int iy = (oy + p - ky) / s;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ox = ox0 + j;
            for (int kx = (ox + p) % s; kx < k; kx += s) {
                const int ix = (ox + p - kx) / s;
                if (ix < 0 || ix >= W) continue;
                acc[j] = fmaf(__ldg(xp + iy * W + ix), __ldg(wp + ky * k + kx), acc[j]);
            }
        }
    }
    *reinterpret_cast<float4 *>(y + ((size_t)bc * Ho + oy) * Wo + ox0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// gx[b,c,iy,ix] = sum_{ky,kx} gy[b,c,iy*s-p+ky, ix*s-p+kx] * w[c,ky,kx]
__global__ void __launch_bounds__(256) dw_deconv_bwd_input_kernel(const float *__restrict__ gy, const float *__restrict__ w,
                                                                 float *__restrict__ gx, int C, int H, int W, int Ho, int Wo,
                                                                 int k, int s, int p)
{
    const int bc = blockIdx.z, c = bc % C;
    const int iy = blockIdx.y;
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    if (ix >= W) return;
    const float *gp = gy + (size_t)bc * Ho * Wo;
    const float *wp = w + (size_t)c * k * k;
    float acc = 0.f;
    for (int ky = 0; ky < k; ++ky) {
        const int oy = iy * s - p + ky;
        if (oy < 0 || oy >= Ho) continue;
        for (int kx = 0; kx < k; ++kx) {
            const int ox = ix * s - p + kx;
            if (ox < 0 || ox >= Wo) continue;
            acc = fmaf(__ldg(gp + oy * Wo + ox), __ldg(wp + ky * k + kx), acc);
        }
    }
    gx[((size_t)bc * H + iy) * W + ix] = acc;
}

// gw[c,ky,kx] = sum_{b,iy,ix} x[b,c,iy,ix] * gy[b,c,iy*s-p+ky, ix*s-p+kx];  grid (k*k, C)
__global__ void __launch_bounds__(256) dw_deconv_bwd_weight_kernel(const float *__restrict__ x, const float *__restrict__ gy,
                                                                  float *__restrict__ gw, int B, int C, int H, int W, int Ho,
                                                                  int Wo, int k, int s, int p)
{
    __shared__ float red[32];
    const int tap = blockIdx.x, c = blockIdx.y;
    const int ky = tap / k, kx = tap - ky * k;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
        const float *xp = x + ((size_t)b * C + c) * H * W;
        const float *gp = gy + ((size_t)b * C + c) * Ho * Wo;
        for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
            const int iy = i / W, ix = i - iy * W;
            const int oy = iy * s - p + ky, ox = ix * s - p + kx;
            if (oy >= 0 && oy < Ho && ox >= 0 && ox < Wo) acc = fmaf(__ldg(xp + i), __ldg(gp + oy * Wo + ox), acc);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) gw[(size_t)c * k * k + tap] = v;
    }
}

// scalar fallback of the fused IDAUp step below for output widths that are not a multiple of 4 (one thread per pixel and channel)
template <int ST>
__global__ void __launch_bounds__(256) idaup_fuse_f16_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                            const float *__restrict__ skip, float *__restrict__ full,
                                                            uint32_t *__restrict__ hi, uint32_t *__restrict__ lo, int C, int H, int W,
                                                            int Cpad, uint32_t *rs)
{
    constexpr int k = 2 * ST, pd = ST / 2;
    __shared__ float tile[64][33];
    const int Ho = H * ST, Wo = W * ST;
    const long long S = (long long)Ho * Wo;
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 64;
    const long long p = p0 + threadIdx.x;
    const int oy = (int)(p / Wo), ox = (int)(p - (long long)oy * Wo);
    for (int i = threadIdx.y; i < 64; i += blockDim.y) {
        const int c = c0 + i;
        float v = 0.f;
        if (c < C && p < S) {
            const float *xp = x + ((size_t)n * C + c) * H * W;
            const float *wp = w + (size_t)c * k * k;
            float acc = 0.f;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const int ky = (oy + pd) % ST + a * ST, iy = (oy + pd - ky) / ST;
                if (iy < 0 || iy >= H) continue;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int kx = (ox + pd) % ST + b * ST, ix = (ox + pd - kx) / ST;
                    if (ix < 0 || ix >= W) continue;
                    acc = fmaf(__ldg(xp + iy * W + ix), __ldg(wp + ky * k + kx), acc);
                }
            }
            v = acc + __ldg(skip + ((size_t)n * C + c) * S + p);
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    const int c = c0 + 2 * threadIdx.x;
    float amax = 0.f;
    for (int i = threadIdx.y; i < 32 && c < Cpad; i += blockDim.y) {
        const long long q = p0 + i;
        if (q >= S) continue;
        const float a = tile[2 * threadIdx.x][i], b = tile[2 * threadIdx.x + 1][i];
        const size_t o = (((size_t)n * S + q) * Cpad + c) >> 1;      // index of the channel PAIR
        uint32_t h, l;
        amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
        f16_split2(a, b, h, l);
        hi[o] = h;
        lo[o] = l;
        reinterpret_cast<float2 *>(full)[o] = make_float2(a, b);
    }
    if (rs) range_commit(rs, amax);
}

// NCHW -> channels-last fp16 operand pairs (+ the unsplit fp32 copy), optionally with one IDAUp step fused in front
// (feature_extraction_dla34.py:380-386):  node_k(up_k(proj_k(layers[i])) + layers[i-1])  -- everything between the two deformable
// convolutions in ONE pass: the depth-wise transposed convolution of the projected map (k = 2 ST, stride ST, padding ST / 2: at
// most 2 x 2 taps per output pixel, same tap order as dw_deconv_fwd_v4_kernel), the skip addition, the change to channels-last and
// the split into the fp16 operand pairs the node's offset convolution reads; the unsplit channels-last copy is what the node's
// deformable gather reads.  Replaces dw_deconv + at::add + ncdhw_to_cl_split (the up-sampled map and the sum never exist in NCHW).
// ST = 0: plain layout change + split of `skip` (side_ncdhw_to_cl_split_f16 for S % 4 == 0).
// Tile = 64 channels x 128 pixels, block (32, 8).  Load phase: a lane reads 4 consecutive pixels of one channel as a 16-byte vector
// (512 contiguous bytes per warp) and stores them to shared memory at the permuted column (px >> 2) + 32 (px & 3) of a row of 129
// floats, so the four scalar stores of a vector are conflict-free.  Store phase: a lane owns 8 channels of one pixel and a warp
// instruction covers the pixels {4 G, 4 G + 4, 4 G + 8, 4 G + 12} + e -- a quarter-warp writes one pixel's 256-byte fp32 row /
// 128-byte hi and lo rows (2-way bank conflicts on the reads).
template <int ST>
__global__ void __launch_bounds__(256) cl_split_tile_f16_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                               const float *__restrict__ skip, float *__restrict__ full,
                                                               uint32_t *__restrict__ hi, uint32_t *__restrict__ lo, int C, int H, int W,
                                                               long long S, int Cpad, uint32_t *rs)
{
    constexpr int T = ST > 0 ? ST : 2;                 // ST = 0 (plain split): the fused branch is compiled out, T only keeps it well-formed
    constexpr int k = 2 * T, pd = T / 2;
    constexpr int kRow = 129;
    __shared__ float tile[64 * kRow];
    const int Wo = ST > 0 ? W * ST : 1;
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 128;
    const int c0 = blockIdx.y * 64;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const long long p = p0 + 4 * tx;                   // S % 4 == 0: a vector is inside the plane or outside as a whole
    int oy = 0, ox0 = 0;
    if constexpr (ST > 0) { oy = (int)(p / Wo); ox0 = (int)(p - (long long)oy * Wo); }      // Wo % 4 == 0: the 4 pixels share a row
    for (int i = ty; i < 64; i += 8) {
        const int c = c0 + i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < C && p < S) {
            v = __ldg(reinterpret_cast<const float4 *>(skip + ((size_t)n * C + c) * S + p));
            if constexpr (ST > 0) {
                const float *xp = x + ((size_t)n * C + c) * H * W;
                const float *wp = w + (size_t)c * k * k;
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                if (T == 2 || T == 4) {
                    // ox0 % 4 == 0 and 4 % T == 0: the tap columns of the four pixels are compile-time offsets from ox0 / T --
                    // pixel j uses kx = (j + pd) % T + b ST and input column ox0 / T + (j + pd) / T - b -- so the two input rows
                    // are read once as a 4-column window and the two weight rows once (same products, same order per pixel)
                    const int xb = ox0 / T;
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const int ky = (oy + pd) % T + a * T, iy = (oy + pd - ky) / T;
                        if (iy < 0 || iy >= H) continue;
                        float xw[4], wr[2 * T];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int ix = xb - 1 + q;
                            xw[q] = (ix >= 0 && ix < W) ? __ldg(xp + iy * W + ix) : 0.f;
                        }
#pragma unroll
                        for (int q = 0; q < 2 * T; ++q) wr[q] = __ldg(wp + ky * k + q);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
#pragma unroll
                            for (int b = 0; b < 2; ++b) {
                                const int kx = (j + pd) % T + b * T, dq = (j + pd) / T - b + 1;      // compile-time
                                const int ix = xb - 1 + dq;
                                if (ix < 0 || ix >= W) continue;
                                acc[j] = fmaf(xw[dq], wr[kx], acc[j]);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const int ky = (oy + pd) % T + a * T, iy = (oy + pd - ky) / T;
                        if (iy < 0 || iy >= H) continue;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ox = ox0 + j;
#pragma unroll
                            for (int b = 0; b < 2; ++b) {
                                const int kx = (ox + pd) % T + b * T, ix = (ox + pd - kx) / T;
                                if (ix < 0 || ix >= W) continue;
                                acc[j] = fmaf(__ldg(xp + iy * W + ix), __ldg(wp + ky * k + kx), acc[j]);
                            }
                        }
                    }
                }
                v.x = acc[0] + v.x; v.y = acc[1] + v.y; v.z = acc[2] + v.z; v.w = acc[3] + v.w;
            }
        }
        float *tr = tile + i * kRow + tx;
        tr[0] = v.x; tr[32] = v.y; tr[64] = v.z; tr[96] = v.w;
    }
    __syncthreads();
    const int o8 = tx & 7, g4 = tx >> 3;               // 8-channel group, which of the instruction's four pixels
    const int c = c0 + 8 * o8;
    float amax = 0.f;
    if (c < Cpad) {
#pragma unroll 1
        for (int it = ty; it < 32; it += 8) {          // it = 4 G' + e: pixels 4 (4 G' + g4) + e
            const int px = 4 * (4 * (it >> 2) + g4) + (it & 3);
            const long long q = p0 + px;
            if (q >= S) continue;
            const float *tc = tile + (8 * o8) * kRow + (px >> 2) + 32 * (px & 3);
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = tc[e * kRow];
            const size_t o = ((size_t)n * S + q) * Cpad + c;
            uint4 h, l;
            f16_split2(v[0], v[1], h.x, l.x); f16_split2(v[2], v[3], h.y, l.y);
            f16_split2(v[4], v[5], h.z, l.z); f16_split2(v[6], v[7], h.w, l.w);
#pragma unroll
            for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(v[e]));
            *reinterpret_cast<uint4 *>(hi + (o >> 1)) = h;
            *reinterpret_cast<uint4 *>(lo + (o >> 1)) = l;
            if (full) {
                *reinterpret_cast<float4 *>(full + o) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(full + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }
    if (rs) range_commit(rs, amax);
}

static int dw_check(int B, int C, int H, int W, int k, int s, int p, int &Ho, int &Wo)
{
    SIDE_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && k > 0 && s > 0 && p >= 0, "dw_deconv: bad shape");
    Ho = (H - 1) * s - 2 * p + k;
    Wo = (W - 1) * s - 2 * p + k;
    SIDE_REQUIRE(Ho > 0 && Wo > 0, "dw_deconv: empty output");
    SIDE_REQUIRE((long long)B * C <= 65535 && Ho <= 65535 && H <= 65535, "dw_deconv: grid too large");
    return SIDE_OK;
}

}  // namespace side

using namespace side;

extern "C" int side_dw_deconv_fwd(const float *x, const float *w, float *y, int B, int C, int H, int W, int k, int stride,
                                  int pad, void *stream)
{
    int Ho, Wo;
    int rc = dw_check(B, C, H, W, k, stride, pad, Ho, Wo);
    if (rc) return rc;
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(w); SIDE_REQUIRE_DEV(y);
    if ((Wo & 3) == 0 && B * C <= 65535 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
        dim3 grid(ceil_div((long long)Ho * (Wo >> 2), 256), B * C);
        cudaStream_t st = (cudaStream_t)stream;
        const bool ida = (k == 2 * stride && pad == stride / 2);
        if (ida && stride == 2) dw_deconv_fwd_v4_kernel<2><<<grid, 256, 0, st>>>(x, w, y, C, H, W, Ho, Wo, k, stride, pad);
        else if (ida && stride == 4) dw_deconv_fwd_v4_kernel<4><<<grid, 256, 0, st>>>(x, w, y, C, H, W, Ho, Wo, k, stride, pad);
        else if (ida && stride == 8) dw_deconv_fwd_v4_kernel<8><<<grid, 256, 0, st>>>(x, w, y, C, H, W, Ho, Wo, k, stride, pad);
        else dw_deconv_fwd_v4_kernel<0><<<grid, 256, 0, st>>>(x, w, y, C, H, W, Ho, Wo, k, stride, pad);
        SIDE_LAUNCH_CHECK("dw_deconv_fwd_v4_kernel");
        return SIDE_OK;
    }
    dim3 grid(ceil_div(Wo, 256), Ho, B * C);
    dw_deconv_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, y, C, H, W, Ho, Wo, k, stride, pad);
    SIDE_LAUNCH_CHECK("dw_deconv_fwd_kernel");
    return SIDE_OK;
}

extern "C" int side_dw_deconv_bwd(const float *x, const float *w, const float *gy, float *gx, float *gw, int B, int C, int H,
                                  int W, int k, int stride, int pad, void *stream)
{
    int Ho, Wo;
    int rc = dw_check(B, C, H, W, k, stride, pad, Ho, Wo);
    if (rc) return rc;
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(w); SIDE_REQUIRE_DEV(gy);
    cudaStream_t st = (cudaStream_t)stream;
    if (gx) {
        SIDE_REQUIRE_DEV(gx);
        dim3 grid(ceil_div(W, 256), H, B * C);
        dw_deconv_bwd_input_kernel<<<grid, 256, 0, st>>>(gy, w, gx, C, H, W, Ho, Wo, k, stride, pad);
        SIDE_LAUNCH_CHECK("dw_deconv_bwd_input_kernel");
    }
    if (gw) {
        SIDE_REQUIRE_DEV(gw);
        dim3 grid(k * k, C);
        dw_deconv_bwd_weight_kernel<<<grid, 256, 0, st>>>(x, gy, gw, B, C, H, W, Ho, Wo, k, stride, pad);
        SIDE_LAUNCH_CHECK("dw_deconv_bwd_weight_kernel");
    }
    return SIDE_OK;
}

extern "C" int side_idaup_fuse_cl_f16(const float *x, const float *w, const float *skip, float *full, void *hi, void *lo, int B,
                                      int C, int H, int W, int stride, int Cpad, void *stream)
{
    SIDE_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && (stride == 2 || stride == 4 || stride == 8),
                 "side_idaup_fuse_cl_f16: bad shape (stride must be 2, 4 or 8)");
    SIDE_REQUIRE(Cpad >= C && Cpad % 8 == 0 && B <= 65535 && (Cpad + 63) / 64 <= 65535,
                 "side_idaup_fuse_cl_f16: bad channel padding / batch (Cpad %% 8 == 0)");
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(w); SIDE_REQUIRE_DEV(skip); SIDE_REQUIRE_DEV(full); SIDE_REQUIRE_DEV(hi); SIDE_REQUIRE_DEV(lo);
    SIDE_REQUIRE(((reinterpret_cast<uintptr_t>(skip) | reinterpret_cast<uintptr_t>(full) | reinterpret_cast<uintptr_t>(hi) |
                   reinterpret_cast<uintptr_t>(lo)) & 15) == 0, "side_idaup_fuse_cl_f16: skip / full / hi / lo must be 16-byte aligned");
    const long long S = (long long)H * stride * W * stride;      // (W * stride) % 4 == 0 always holds: stride is even ... times W
    if ((W * stride) % 4 != 0) {
        const dim3 g1((unsigned)((S + 31) / 32), (unsigned)((Cpad + 63) / 64), (unsigned)B), b1(32, 8);
        cudaStream_t s1 = (cudaStream_t)stream;
        uint32_t *h1 = reinterpret_cast<uint32_t *>(hi), *l1 = reinterpret_cast<uint32_t *>(lo);
        if (stride == 2) idaup_fuse_f16_kernel<2><<<g1, b1, 0, s1>>>(x, w, skip, full, h1, l1, C, H, W, Cpad, range_slot_next());
        else if (stride == 4) idaup_fuse_f16_kernel<4><<<g1, b1, 0, s1>>>(x, w, skip, full, h1, l1, C, H, W, Cpad, range_slot_next());
        else idaup_fuse_f16_kernel<8><<<g1, b1, 0, s1>>>(x, w, skip, full, h1, l1, C, H, W, Cpad, range_slot_next());
        SIDE_LAUNCH_CHECK("idaup_fuse_f16_kernel");
        return SIDE_OK;
    }
    const dim3 grid((unsigned)((S + 127) / 128), (unsigned)((Cpad + 63) / 64), (unsigned)B), block(32, 8);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *h = reinterpret_cast<uint32_t *>(hi), *l = reinterpret_cast<uint32_t *>(lo);
    if (stride == 2) cl_split_tile_f16_kernel<2><<<grid, block, 0, st>>>(x, w, skip, full, h, l, C, H, W, S, Cpad, range_slot_next());
    else if (stride == 4) cl_split_tile_f16_kernel<4><<<grid, block, 0, st>>>(x, w, skip, full, h, l, C, H, W, S, Cpad, range_slot_next());
    else cl_split_tile_f16_kernel<8><<<grid, block, 0, st>>>(x, w, skip, full, h, l, C, H, W, S, Cpad, range_slot_next());
    SIDE_LAUNCH_CHECK("cl_split_tile_f16_kernel<fused>");
    return SIDE_OK;
}

// side_ncdhw_to_cl_split_f16 without a scale, S % 4 == 0, Cpad % 8 == 0, aligned pointers: the tiled kernel above (aggr_ops.cu)
namespace side {
int launch_cl_split_tile_f16(const float *x, float *full, void *hi, void *lo, int N, int C, long long S, int Cpad, cudaStream_t st)
{
    const dim3 grid((unsigned)((S + 127) / 128), (unsigned)((Cpad + 63) / 64), (unsigned)N), block(32, 8);
    cl_split_tile_f16_kernel<0><<<grid, block, 0, st>>>(nullptr, nullptr, x, full, reinterpret_cast<uint32_t *>(hi),
                                                        reinterpret_cast<uint32_t *>(lo), C, 0, 0, S, Cpad, range_slot_next());
    SIDE_LAUNCH_CHECK("cl_split_tile_f16_kernel");
    return SIDE_OK;
}
}  // namespace side
