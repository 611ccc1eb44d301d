// stem_conv.cu -- the DLA-34 stem: base_layer 7x7 (3 -> 16), level0 3x3 (16 -> 16) and
// level1 3x3 stride 2 (16 -> 32) at full resolution, each followed by BatchNorm + ReLU (feature_extraction_dla34.py
// :281-293 of the reference, `_make_conv_level`).  16-32 channels give the tensor cores nothing to chew on (an
// M128 x N16 tf32 MMA is bound by its A-operand fetch), so these are direct fp32 SIMT convolutions: a CTA stages the input
// halo of a 64 x 8 output tile and the whole filter bank in shared memory; a thread owns 2 output pixels (columns lane and
// lane + 32 of one row) and ALL output channels, so every input value it loads feeds Cout FMAs and every (broadcast) 16-byte
// weight load feeds 8.  Shared-memory reads are conflict-free: lanes read consecutive words (stride-2 layers stage the even
// and the odd input columns in separate half rows).
// Eval-mode BatchNorm and ReLU are folded into the store.  cuDNN spends 3-4x longer on these three layers (its fp32
// kernels are tuned for wide channels) plus separate BN / ReLU passes.
#include "tc_common.cuh"

namespace side {

constexpr int kStemTW = 64, kStemTH = 8;     // output tile; 256 threads, 2 pixels each

// S2D: instead of NCHW fp32 the result leaves as fp16 (hi, lo * 2^11) operand pairs in the 2x2 space-to-depth channels-last layout
// [B, Ho/2, Wo/2, 4 * COUT], channel = (dy * 2 + dx) * COUT + o -- the input format of the next two stem layers when they run as
// 3x3 block convolutions on the tensor cores (DLA._stem_tc).  y = hi array, y2 = lo array.
template <int CIN, int COUT, int K, int S, int CCHUNK, bool S2D = false>
__global__ void __launch_bounds__(256) stem_conv_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                        const float *__restrict__ scale, const float *__restrict__ shift,
                                                        float *__restrict__ y, int H, int W, int Ho, int Wo, int relu,
                                                        void *__restrict__ y2 = nullptr, uint32_t *rs = nullptr)
{
    constexpr int P = (K - 1) / 2;
    constexpr int IW = kStemTW * S + K - 1, IH = kStemTH * S + K - 1;
    constexpr int HALF = (IW + 1) / 2;                           // S == 2: even columns at [0, HALF), odd ones at [HALF, 2 HALF)
    constexpr int IWP = (S == 2 ? 2 * HALF : IW) | 1;            // odd row pitch
    extern __shared__ __align__(16) float sm[];
    float *ws = sm;                                              // [CIN * K * K][COUT]
    float *in_s = sm + CIN * K * K * COUT;                       // [CCHUNK][IH][IWP]
    const int b = blockIdx.z, oy0 = blockIdx.y * kStemTH, ox0 = blockIdx.x * kStemTW;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;   // output columns tx and tx + 32 of tile row ty
    for (int i = tid; i < CIN * K * K * COUT; i += 256) {
        const int o = i % COUT, ck = i / COUT;                   // ws[ck][o] = w[o][ck]   (w is [COUT][CIN][K][K])
        ws[i] = __ldg(w + (size_t)o * CIN * K * K + ck);
    }
    // accumulators as channel PAIRS: fma.rn.f32x2 (FFMA2) does two of them per issue slot -- the layer is bound by instruction
    // issue (1600 FFMA + 294 LDS per input channel and thread), not by the FP32 pipe; same IEEE results
    float2 acc0[COUT / 2], acc1[COUT / 2];
#pragma unroll
    for (int o = 0; o < COUT / 2; ++o) acc0[o] = acc1[o] = make_float2(0.f, 0.f);
    const float *xb = x + (size_t)b * CIN * H * W;
    const int iy0 = oy0 * S - P, ix0 = ox0 * S - P;
    for (int c0 = 0; c0 < CIN; c0 += CCHUNK) {
        __syncthreads();
        for (int i = tid; i < CCHUNK * IH * IW; i += 256) {
            const int xx = i % IW, r = i / IW, yy = r % IH, c = r / IH;
            const int gy = iy0 + yy, gx = ix0 + xx;
            // asynchronous 4-byte copies (zero fill outside the image): all of a thread's loads are in flight at once instead
            // of one load -> store round trip per element, which was 70 % of the stride-2 layer's time
            const bool ok = c0 + c < CIN && gy >= 0 && gy < H && gx >= 0 && gx < W;
            const float *src = ok ? xb + ((size_t)(c0 + c) * H + gy) * W + gx : xb;
            const uint32_t dst = smem_u32(&in_s[(c * IH + yy) * IWP + (S == 2 ? (xx & 1) * HALF + (xx >> 1) : xx)]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(ok ? 4 : 0) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
#pragma unroll 1
        for (int c = 0; c < CCHUNK && c0 + c < CIN; ++c) {
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                const float *row = in_s + (c * IH + ty * S + ky) * IWP + tx;
                const float *wr = ws + ((c0 + c) * K * K + ky * K) * COUT;
#pragma unroll
                for (int kx = 0; kx < K; ++kx) {
                    // input column tx * S + kx (and 32 * S further right for the second pixel)
                    const int ci = S == 2 ? (kx & 1) * HALF + (kx >> 1) : kx;
                    const float v0 = row[ci], v1 = row[ci + 32];
                    const float2 v00 = make_float2(v0, v0), v11 = make_float2(v1, v1);
#pragma unroll
                    for (int o4 = 0; o4 < COUT / 4; ++o4) {
                        const float4 wv = *reinterpret_cast<const float4 *>(wr + kx * COUT + 4 * o4);
                        const float2 wa = make_float2(wv.x, wv.y), wb = make_float2(wv.z, wv.w);
                        acc0[2 * o4] = __ffma2_rn(v00, wa, acc0[2 * o4]);         acc1[2 * o4] = __ffma2_rn(v11, wa, acc1[2 * o4]);
                        acc0[2 * o4 + 1] = __ffma2_rn(v00, wb, acc0[2 * o4 + 1]); acc1[2 * o4 + 1] = __ffma2_rn(v11, wb, acc1[2 * o4 + 1]);
                    }
                }
            }
        }
    }
    const int oy = oy0 + ty, ox = ox0 + tx;
    if (S2D) {
        float amax = 0.f;
        if (oy < Ho) {
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                const int xo = ox + 32 * px;
                if (xo >= Wo) continue;
                const float2 *acc = px ? acc1 : acc0;
                uint32_t hh[COUT / 2], ll[COUT / 2];
#pragma unroll
                for (int o = 0; o < COUT; o += 2) {
                    float a = fmaf(acc[o / 2].x, scale ? __ldg(scale + o) : 1.f, shift ? __ldg(shift + o) : 0.f);
                    float c = fmaf(acc[o / 2].y, scale ? __ldg(scale + o + 1) : 1.f, shift ? __ldg(shift + o + 1) : 0.f);
                    if (relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
                    amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(c)));
                    f16_split2(a, c, hh[o / 2], ll[o / 2]);
                }
                const size_t off = ((((size_t)b * (Ho >> 1) + (oy >> 1)) * (Wo >> 1) + (xo >> 1)) * 4 + ((oy & 1) * 2 + (xo & 1))) * COUT;
                uint4 *hp = reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(y) + off);
                uint4 *lp = reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(y2) + off);
#pragma unroll
                for (int q = 0; q < COUT / 8; ++q) {
                    hp[q] = make_uint4(hh[4 * q], hh[4 * q + 1], hh[4 * q + 2], hh[4 * q + 3]);
                    lp[q] = make_uint4(ll[4 * q], ll[4 * q + 1], ll[4 * q + 2], ll[4 * q + 3]);
                }
            }
        }
        if (rs) range_commit(rs, amax);
        return;
    }
    if (oy < Ho && ox < Wo) {
        float *yp = y + ((size_t)b * COUT * Ho + oy) * Wo + ox;
        const bool two = ox + 32 < Wo;
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            const float sc = scale ? __ldg(scale + o) : 1.f, sh = shift ? __ldg(shift + o) : 0.f;
            float a = fmaf(o & 1 ? acc0[o / 2].y : acc0[o / 2].x, sc, sh), c = fmaf(o & 1 ? acc1[o / 2].y : acc1[o / 2].x, sc, sh);
            if (relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
            yp[(size_t)o * Ho * Wo] = a;                         // a warp writes 128 contiguous bytes per channel and half tile
            if (two) yp[(size_t)o * Ho * Wo + 32] = c;
        }
    }
}

template <int CIN, int COUT, int K, int S, int CCHUNK>
static int launch_stem(const float *x, const float *w, const float *scale, const float *shift, float *y, int B, int H, int W,
                       int relu, cudaStream_t st)
{
    constexpr int P = (K - 1) / 2;
    const int Ho = (H + 2 * P - K) / S + 1, Wo = (W + 2 * P - K) / S + 1;
    constexpr int IW = kStemTW * S + K - 1, IH = kStemTH * S + K - 1, IWP = (S == 2 ? 2 * ((IW + 1) / 2) : IW) | 1;
    const size_t smem = sizeof(float) * ((size_t)CIN * K * K * COUT + (size_t)CCHUNK * IH * IWP);
    int rc = set_smem_attr((const void *)stem_conv_kernel<CIN, COUT, K, S, CCHUNK>, smem);
    if (rc) return rc;
    dim3 grid(ceil_div(Wo, kStemTW), ceil_div(Ho, kStemTH), B);
    stem_conv_kernel<CIN, COUT, K, S, CCHUNK><<<grid, 256, smem, st>>>(x, w, scale, shift, y, H, W, Ho, Wo, relu);
    SIDE_LAUNCH_CHECK("stem_conv_kernel");
    return SIDE_OK;
}

}  // namespace side

using namespace side;

extern "C" int side_stem_conv_fwd(const float *x, const float *w, const float *scale, const float *shift, float *y, int B, int Cin,
                                  int H, int W, int Cout, int k, int stride, int relu, void *stream)
{
    SIDE_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "side_stem_conv_fwd: bad shape");
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(w); SIDE_REQUIRE_DEV(y);
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 3 && Cout == 16 && k == 7 && stride == 1) return launch_stem<3, 16, 7, 1, 3>(x, w, scale, shift, y, B, H, W, relu, st);
    if (Cin == 16 && Cout == 16 && k == 3 && stride == 1) return launch_stem<16, 16, 3, 1, 16>(x, w, scale, shift, y, B, H, W, relu, st);
    if (Cin == 16 && Cout == 32 && k == 3 && stride == 2) return launch_stem<16, 32, 3, 2, 8>(x, w, scale, shift, y, B, H, W, relu, st);
    set_error("side_stem_conv_fwd: only the DLA-34 stem shapes are built (3->16 k7 s1, 16->16 k3 s1, 16->32 k3 s2), got %d->%d k%d s%d",
              Cin, Cout, k, stride);
    return SIDE_ERR_UNSUPPORTED;
}

/* base_layer (3 -> 16, 7x7, stride 1, folded BatchNorm + ReLU) with the result written as fp16 operand pairs in the 2x2
 * space-to-depth channels-last layout [B, H/2, W/2, 64] (channel = (dy * 2 + dx) * 16 + o): what DLA level0 / level1 read when they
 * run as 3x3 block convolutions on side_conv3d_tc_fwd_f16.  H, W even. */
extern "C" int side_stem_conv_fwd_s2d(const float *x, const float *w, const float *scale, const float *shift, void *y_hi, void *y_lo,
                                      int B, int H, int W, int relu, void *stream)
{
    SIDE_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "side_stem_conv_fwd_s2d: bad shape");
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(w); SIDE_REQUIRE_DEV(y_hi); SIDE_REQUIRE_DEV(y_lo);
    constexpr int CIN = 3, COUT = 16, K = 7, S = 1, CCHUNK = 3;
    constexpr int IW = kStemTW * S + K - 1, IH = kStemTH * S + K - 1, IWP = IW | 1;
    const size_t smem = sizeof(float) * ((size_t)CIN * K * K * COUT + (size_t)CCHUNK * IH * IWP);
    int rc = set_smem_attr((const void *)stem_conv_kernel<CIN, COUT, K, S, CCHUNK, true>, smem);
    if (rc) return rc;
    dim3 grid(ceil_div(W, kStemTW), ceil_div(H, kStemTH), B);
    stem_conv_kernel<CIN, COUT, K, S, CCHUNK, true><<<grid, 256, smem, (cudaStream_t)stream>>>(
        x, w, scale, shift, reinterpret_cast<float *>(y_hi), H, W, H, W, relu, y_lo, range_slot_next());
    SIDE_LAUNCH_CHECK("stem_conv_kernel<s2d>");
    return SIDE_OK;
}
