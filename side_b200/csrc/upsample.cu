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

// IDAUp step fused (feature_extraction_dla34.py:380-386):  node_k(up_k(proj_k(layers[i])) + layers[i-1])  -- everything between
// the two deformable convolutions in ONE pass: the depth-wise transposed convolution of the projected map (k = 2 ST, stride ST,
// padding ST / 2: at most 2 x 2 taps per output pixel, same tap order as dw_deconv_fwd_v4_kernel), the skip addition, the change to
// channels-last and the split into the fp16 operand pairs the node's offset convolution reads; the unsplit channels-last copy is
// what the node's deformable gather reads.  Replaces dw_deconv + at::add + ncdhw_to_cl_split (the up-sampled map and the sum
// never exist in NCHW).  Block (32, 8): 64 channels x 32 pixels through a shared-memory transpose, a lane owns two adjacent
// channels so every warp store is 128 contiguous bytes of one pixel's row.
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
    SIDE_REQUIRE(Cpad >= C && Cpad % 2 == 0 && B <= 65535 && (Cpad + 63) / 64 <= 65535, "side_idaup_fuse_cl_f16: bad channel padding / batch");
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(w); SIDE_REQUIRE_DEV(skip); SIDE_REQUIRE_DEV(full); SIDE_REQUIRE_DEV(hi); SIDE_REQUIRE_DEV(lo);
    const long long S = (long long)H * stride * W * stride;
    const dim3 grid((unsigned)((S + 31) / 32), (unsigned)((Cpad + 63) / 64), (unsigned)B), block(32, 8);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *h = reinterpret_cast<uint32_t *>(hi), *l = reinterpret_cast<uint32_t *>(lo);
    if (stride == 2) idaup_fuse_f16_kernel<2><<<grid, block, 0, st>>>(x, w, skip, full, h, l, C, H, W, Cpad, range_slot_next());
    else if (stride == 4) idaup_fuse_f16_kernel<4><<<grid, block, 0, st>>>(x, w, skip, full, h, l, C, H, W, Cpad, range_slot_next());
    else idaup_fuse_f16_kernel<8><<<grid, block, 0, st>>>(x, w, skip, full, h, l, C, H, W, Cpad, range_slot_next());
    SIDE_LAUNCH_CHECK("idaup_fuse_f16_kernel");
    return SIDE_OK;
}
