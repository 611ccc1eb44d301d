// conv_wgrad.cu -- weight gradient of the plain convolutions on tcgen05 (training side).
//
// Reference: the reference trains this network (modules/stereoTrainer.py:254-319); every nn.Conv2d / nn.Conv3d of the DLA-34 levels,
// the heads and the 3-D aggregation network (stereo_network_old.py:139-171, feature_extraction_dla34.py:31-95) gets its weight
// gradient from cuDNN's fp32 SIMT wgrad kernels (28 ms per step for the seven Conv3d layers alone at 16 RoIs).  Here
//     gW[o][tap][c] = sum over (sample, output voxel) of gy[n][o][v] * x[n][v * stride + tap - pad][c]
// runs as the split-K tensor-core GEMM of the DCN weight gradient (dcn_gw_tc.cu: gy is a K-major A operand as it lies in memory,
// the pixel-major patch matrix is read in place as an MN-major B operand, 3xFP16 pairs with a power-of-two range scale on gy):
//   dcn_gw_tc_split_gy        fp32 gy -> scaled fp16 pairs (once per call);
//   dcn_gw_tc_kernel, fused mode: a k-block = a box of 64 output voxels; the B operand of tap t is the SAME box of the saved
//                             input pairs shifted by the tap (5-D TMA, traversal stride = convolution stride, out-of-bound fill =
//                             the zero padding) -- the patch matrix never exists;
//   conv_im2col_pairs_kernel + dcn_gw_tc_kernel: fallback for output maps that do not tile into 64-voxel boxes: the pairs are
//                             copied into the patch matrix [sample * voxel][tap * Cq + c] in chunks of samples.
#include <algorithm>

#include "tc_common.cuh"

namespace side {

bool dcn_gw_tc_supported(int Cout, int Cin, int KK, int P, long long rows, bool partial_ok = false);
size_t dcn_gw_tc_gy_halves(int B, int Cout, int P);
int dcn_gw_tc_split_gy(const float *gy, void *gy_pairs, int B, int Cout, int P, cudaStream_t st);
int dcn_gw_tc_run(const void *gy_pairs, const void *col_hi, const void *col_lo, float *gw, int B, int b0, int nb, int Cout, int Kp,
                  int P, cudaStream_t st);
bool dcn_gw_box(int Do, int Ho, int Wo, int &bw, int &bh, int &bd);
int dcn_gw_tc_split_gy_box(const float *gy, void *gy_pairs, int N, int Cout, int Do, int Ho, int Wo, int bw, int bh, int bd,
                           cudaStream_t st);
int dcn_gw_tc_run_fused(const void *gy_pairs, const void *x_hi, const void *x_lo, float *gw, int N, int Cout, int Cp, int D, int H, int W,
                        int Do, int Ho, int Wo, int kd, int kh, int kw, int stride, cudaStream_t st);

struct Im2colParams {
    const uint4 *x_hi, *x_lo;          // [N, D, H, W, Cp] halves, 8 per uint4
    uint4 *col_hi, *col_lo;            // [nb * P][taps * Cp]
    int n0, nb, D, H, W, Cp8, Cq8, Do, Ho, Wo, kd, kh, kw, stride;    // Cq8: row chunks per tap (Cp rounded up to 64 channels)
};

__global__ void __launch_bounds__(256) conv_im2col_pairs_kernel(Im2colParams p)
{
    const int taps = p.kd * p.kh * p.kw;
    const long long P = (long long)p.Do * p.Ho * p.Wo;
    const long long total = (long long)p.nb * P * taps * p.Cq8;
    const int pd = (p.kd - 1) / 2, ph = (p.kh - 1) / 2, pw = (p.kw - 1) / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % p.Cq8);
        long long r = i / p.Cq8;
        const int tap = (int)(r % taps);
        r /= taps;                                   // (sample-local, voxel)
        const int wo = (int)(r % p.Wo);
        long long r2 = r / p.Wo;
        const int ho = (int)(r2 % p.Ho);
        r2 /= p.Ho;
        const int d_o = (int)(r2 % p.Do), nl = (int)(r2 / p.Do);
        const int tw = tap % p.kw, th = (tap / p.kw) % p.kh, td = tap / (p.kw * p.kh);
        const int di = d_o + td - pd, hi = ho * p.stride + th - ph, wi = wo * p.stride + tw - pw;
        uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
        if (c8 < p.Cp8 && di >= 0 && di < p.D && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) {
            const size_t src = ((((size_t)(p.n0 + nl) * p.D + di) * p.H + hi) * p.W + wi) * p.Cp8 + c8;
            a = __ldg(p.x_hi + src);
            b = __ldg(p.x_lo + src);
        }
        __stcs(p.col_hi + i, a);
        __stcs(p.col_lo + i, b);
    }
}

// max |x| -> power-of-two range scale for the fp16 pairs of a back-propagated gradient (input-gradient convolution): s brings the
// largest element into [2^10, 2^11); s goes to scale_out[0 .. n_scale), 1 / s to inv_out[0 .. n_inv); all-zero / non-finite -> 1
__global__ void __launch_bounds__(256) range_absmax_kernel(const float *__restrict__ x, long long n, uint32_t *__restrict__ bits)
{
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(__ldg(x + i)));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));
}
__global__ void range_scale_fill_kernel(const uint32_t *__restrict__ bits, float *__restrict__ scale_out, int n_scale,
                                        float *__restrict__ inv_out, int n_inv)
{
    const uint32_t b = *bits;
    int e = (int)(b >> 23) - 127;
    if (b < 0x00800000u || b >= 0x7F800000u) e = 10;
    const int sh = max(-126, min(126, 10 - e));
    const float s = __uint_as_float((uint32_t)(127 + sh) << 23), inv = __uint_as_float((uint32_t)(127 - sh) << 23);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max(n_scale, n_inv); i += gridDim.x * blockDim.x) {
        if (i < n_scale) scale_out[i] = s;
        if (i < n_inv) inv_out[i] = inv;
    }
}

constexpr size_t kWgradColBudget = (size_t)768 << 20;      // bytes of patch matrix (hi + lo) per chunk of samples

static int wgrad_chunk(int N, long long P, int Kp)
{
    const size_t per = (size_t)P * Kp * 4;                  // hi + lo halves of one sample
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)N, kWgradColBudget / std::max<size_t>(per, 1)));
}

}  // namespace side

using namespace side;

static bool wgrad_shape(int N, int D, int H, int W, int Cp, int Cout, int kd, int kh, int kw, int stride, int &Do, int &Ho, int &Wo)
{
    if (N <= 0 || D <= 0 || H <= 0 || W <= 0 || Cp <= 0 || Cout <= 0 || Cp % 8 != 0) return false;
    if (!((kd == 1 || kd == 3) && (kh == 1 || kh == 3) && kh == kw && (stride == 1 || stride == 2))) return false;
    Do = D; Ho = (H + 2 * ((kh - 1) / 2) - kh) / stride + 1; Wo = (W + 2 * ((kw - 1) / 2) - kw) / stride + 1;
    return Ho > 0 && Wo > 0;
}

extern "C" size_t side_conv_wgrad_tc_ws_bytes(int N, int D, int H, int W, int Cp, int Cout, int kd, int kh, int kw, int stride)
{
    int Do, Ho, Wo;
    if (!wgrad_shape(N, D, H, W, Cp, Cout, kd, kh, kw, stride, Do, Ho, Wo)) return 0;
    const long long P = (long long)Do * Ho * Wo;
    const int Kp = kd * kh * kw * ((Cp + 63) / 64 * 64);
    const size_t gy = (2 * dcn_gw_tc_gy_halves(N, Cout, (int)P) + 256 + 255) & ~(size_t)255;
    return gy + (size_t)wgrad_chunk(N, P, Kp) * P * Kp * 4 + 256;
}

extern "C" int side_conv_wgrad_tc(const void *x_hi, const void *x_lo, const float *gy, float *gw, int N, int D, int H, int W, int Cp,
                                  int Cout, int kd, int kh, int kw, int stride, void *ws, size_t ws_bytes, void *stream)
{
    int Do, Ho, Wo;
    SIDE_REQUIRE(wgrad_shape(N, D, H, W, Cp, Cout, kd, kh, kw, stride, Do, Ho, Wo),
                 "side_conv_wgrad_tc: kernels 1x1x1 / 1x3x3 / 3x3x3 with padding (k-1)/2, stride 1 or 2, Cp %% 8 == 0");
    const long long P = (long long)Do * Ho * Wo;
    const int taps = kd * kh * kw, Cq = (Cp + 63) / 64 * 64, Kp = taps * Cq;
    if (!(P < (1ll << 31) && dcn_gw_tc_supported(Cout, Cq, taps, (int)P, (long long)N * P, true))) {
        set_error("side_conv_wgrad_tc: needs Do*Ho*Wo %% 8 == 0 and Cout %% 8 == 0 (got P=%lld, Cout=%d)", P, Cout);
        return SIDE_ERR_UNSUPPORTED;
    }
    SIDE_REQUIRE_DEV(x_hi); SIDE_REQUIRE_DEV(x_lo); SIDE_REQUIRE_DEV(gy); SIDE_REQUIRE_DEV(gw);
    if (ws == nullptr || ws_bytes < side_conv_wgrad_tc_ws_bytes(N, D, H, W, Cp, Cout, kd, kh, kw, stride) || !is_device_ptr(ws) ||
        (reinterpret_cast<uintptr_t>(ws) & 255)) {
        set_error("side_conv_wgrad_tc: needs side_conv_wgrad_tc_ws_bytes(...) bytes of 256-byte aligned device workspace");
        return SIDE_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char *w8 = reinterpret_cast<unsigned char *>(ws);
    void *gy_pairs = w8;
    const size_t gy_bytes = (2 * dcn_gw_tc_gy_halves(N, Cout, (int)P) + 256 + 255) & ~(size_t)255;
    const int nbmax = wgrad_chunk(N, P, Kp);
    __half *col_hi = reinterpret_cast<__half *>(w8 + gy_bytes), *col_lo = col_hi + (size_t)nbmax * P * Kp;
    SIDE_CUDA(cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)Cout * Kp, st));
    int rc;
    // first choice: the input is read in place through shifted TMA boxes (no patch matrix) against gy pairs written box-major;
    // maps that do not tile into 64-voxel boxes (12 x 40 at D = 1) take the patch-matrix path below
    int bw, bh, bd;
    if (dcn_gw_box(Do, Ho, Wo, bw, bh, bd)) {
        if ((rc = dcn_gw_tc_split_gy_box(gy, gy_pairs, N, Cout, Do, Ho, Wo, bw, bh, bd, st))) return rc;
        rc = dcn_gw_tc_run_fused(gy_pairs, x_hi, x_lo, gw, N, Cout, Cp, D, H, W, Do, Ho, Wo, kd, kh, kw, stride, st);
        if (rc != SIDE_ERR_UNSUPPORTED) return rc;
    }
    if ((rc = dcn_gw_tc_split_gy(gy, gy_pairs, N, Cout, (int)P, st))) return rc;
    for (int b0 = 0; b0 < N; b0 += nbmax) {
        const int nb = std::min(nbmax, N - b0);
        Im2colParams ip{reinterpret_cast<const uint4 *>(x_hi), reinterpret_cast<const uint4 *>(x_lo), reinterpret_cast<uint4 *>(col_hi),
                        reinterpret_cast<uint4 *>(col_lo), b0, nb, D, H, W, Cp / 8, Cq / 8, Do, Ho, Wo, kd, kh, kw, stride};
        const long long total = (long long)nb * P * taps * (Cq / 8);
        conv_im2col_pairs_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 32), 256, 0, st>>>(ip);
        SIDE_LAUNCH_CHECK("conv_im2col_pairs_kernel");
        if ((rc = dcn_gw_tc_run(gy_pairs, col_hi, col_lo, gw, N, b0, nb, Cout, Kp, (int)P, st))) return rc;
    }
    return SIDE_OK;
}

extern "C" int side_pow2_range_scale(const float *x, long long n, float *scale_out, int n_scale, float *inv_out, int n_inv,
                                     void *scratch_word, void *stream)
{
    SIDE_REQUIRE(n > 0 && n_scale >= 0 && n_inv >= 0, "side_pow2_range_scale: bad sizes");
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(scratch_word);
    if (n_scale) SIDE_REQUIRE_DEV(scale_out);
    if (n_inv) SIDE_REQUIRE_DEV(inv_out);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *bits = reinterpret_cast<uint32_t *>(scratch_word);
    SIDE_CUDA(cudaMemsetAsync(bits, 0, sizeof(uint32_t), st));
    range_absmax_kernel<<<(unsigned)std::min<long long>((n + 1023) / 1024, 148 * 8), 256, 0, st>>>(x, n, bits);
    SIDE_LAUNCH_CHECK("range_absmax_kernel");
    range_scale_fill_kernel<<<(unsigned)std::max(1, std::min(64, (std::max(n_scale, n_inv) + 255) / 256)), 256, 0, st>>>(bits, scale_out,
                                                                                                                    n_scale, inv_out, n_inv);
    SIDE_LAUNCH_CHECK("range_scale_fill_kernel");
    return SIDE_OK;
}
