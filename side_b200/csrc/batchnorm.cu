// batchnorm.cu -- training-mode BatchNorm forward / backward (training side of SURVEY.md 8f rows F1 / F4; BASELINE config #5).
//
// Reference: every nn.BatchNorm2d / nn.BatchNorm3d of the DLA-34 levels and of the aggregation network in train() mode
// (feature_extraction_dla34.py:31-95, stereo_network_old.py:139-171) -- cuDNN's bn_fw_tr / bn_bw_1C11 kernels on NCHW tensors,
// 18 ms of a 97 ms training step at 5-9x the time of one pass over the data.  Here: NCHW / NCDHW in place, no layout change.
//   bn_stats_kernel      per-channel sum and sum of squares in double, grid = (channel, slice of the (n, s) range);
//   bn_finalize_kernel   mean, biased variance, 1/sqrt(var + eps), running statistics (unbiased variance, momentum);
//   bn_apply_kernel      y = gamma * (x - mean) * invstd + beta, 16-byte accesses;
//   bn_bwd_stats_kernel  sum(gy) and sum(gy * (x - mean)) per channel (double);
//   bn_bwd_apply_kernel  gx = (gy - sum_gy / M - (x - mean) * invstd^2 * sum_gy_xmu / M) * invstd * gamma   (ATen's formula).
#include <algorithm>

#include "common.cuh"

namespace side {

constexpr int kBnThreads = 256;

// work split of one channel's N * S elements over `splits` CTAs: CTA j takes chunks j, j + splits, ... of kBnChunk elements
constexpr int kBnChunk = 4096;

__device__ __forceinline__ double bn_block_sum(double v, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kBnThreads / 32; ++w) t += red[w];
    return t;
}

// BWD: a = gy, b = gy * (x - mean);  FWD: a = x, b = x * x
template <bool BWD>
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float *__restrict__ x, const float *__restrict__ gy,
                                                             const float *__restrict__ mean, double *__restrict__ partial, int N,
                                                             int C, long long S, int splits)
{
    __shared__ double red[kBnThreads / 32];
    const int c = blockIdx.x, j = blockIdx.y;
    const long long chunks_per_n = (S + kBnChunk - 1) / kBnChunk, chunks = chunks_per_n * N;
    const float mu = BWD ? __ldg(mean + c) : 0.f;
    double sa = 0.0, sb = 0.0;
    for (long long ch = j; ch < chunks; ch += splits) {
        const long long n = ch / chunks_per_n, s0 = (ch - n * chunks_per_n) * kBnChunk;
        const long long len = min((long long)kBnChunk, S - s0);
        const size_t base = ((size_t)n * C + c) * S + s0;
        float fa = 0.f, fb = 0.f;                         // <= 16 elements per thread per chunk: fp32 inside, double across
        if (((base | (size_t)len) & 3) == 0) {
            for (long long i = 4 * threadIdx.x; i < len; i += 4 * kBnThreads) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(x + base + i));
                if (BWD) {
                    const float4 g = __ldg(reinterpret_cast<const float4 *>(gy + base + i));
                    fa += (g.x + g.y) + (g.z + g.w);
                    fb += g.x * (v.x - mu) + g.y * (v.y - mu) + g.z * (v.z - mu) + g.w * (v.w - mu);
                } else {
                    fa += (v.x + v.y) + (v.z + v.w);
                    fb += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
                }
            }
        } else {
            for (long long i = threadIdx.x; i < len; i += kBnThreads) {
                const float v = __ldg(x + base + i);
                if (BWD) { const float g = __ldg(gy + base + i); fa += g; fb += g * (v - mu); }
                else { fa += v; fb += v * v; }
            }
        }
        sa += (double)fa; sb += (double)fb;
    }
    sa = bn_block_sum(sa, red);
    sb = bn_block_sum(sb, red);
    if (threadIdx.x == 0) {
        partial[((size_t)c * splits + j) * 2] = sa;
        partial[((size_t)c * splits + j) * 2 + 1] = sb;
    }
}

__global__ void bn_finalize_kernel(const double *__restrict__ partial, int C, int splits, double M, float eps, float momentum,
                                   float *__restrict__ save_mean, float *__restrict__ save_invstd, float *__restrict__ running_mean,
                                   float *__restrict__ running_var)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int j = 0; j < splits; ++j) { s += partial[((size_t)c * splits + j) * 2]; q += partial[((size_t)c * splits + j) * 2 + 1]; }
    const double mean = s / M;
    double var = q / M - mean * mean;                    // double: no cancellation problem at fp32 data
    if (var < 0.0) var = 0.0;
    save_mean[c] = (float)mean;
    save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(M > 1.0 ? var * M / (M - 1.0) : var);
}

// sums of the backward reduced per channel: out[c] = {sum_gy, sum_gy_xmu}; also the parameter gradients
__global__ void bn_bwd_finalize_kernel(const double *__restrict__ partial, int C, int splits, const float *__restrict__ invstd,
                                       float *__restrict__ sums, float *__restrict__ ggamma, float *__restrict__ gbeta)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int j = 0; j < splits; ++j) { s += partial[((size_t)c * splits + j) * 2]; q += partial[((size_t)c * splits + j) * 2 + 1]; }
    sums[2 * c] = (float)s; sums[2 * c + 1] = (float)q;
    if (ggamma) ggamma[c] = (float)(q * (double)invstd[c]);
    if (gbeta) gbeta[c] = (float)s;
}

// FWD: y = gamma * (x - mean) * invstd + beta.  BWD: gx = (gy - s0 / M - (x - mean) * invstd^2 * s1 / M) * invstd * gamma
template <bool BWD>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const float *__restrict__ x, const float *__restrict__ gy,
                                                             const float *__restrict__ gamma, const float *__restrict__ beta,
                                                             const float *__restrict__ mean, const float *__restrict__ invstd,
                                                             const float *__restrict__ sums, float inv_m, float *__restrict__ out,
                                                             int C, long long S)
{
    const long long plane = blockIdx.y;                  // n * C + c
    const int c = (int)(plane % C);
    const float mu = __ldg(mean + c), is = __ldg(invstd + c), g = gamma ? __ldg(gamma + c) : 1.f;
    float k0, k1, k2;
    if (BWD) {
        // gx = gy * (is g) - [(s0 / M) is g] - (x - mu) * [is^3 g s1 / M]
        k0 = is * g;
        k1 = __ldg(sums + 2 * c) * inv_m * k0;
        k2 = __ldg(sums + 2 * c + 1) * inv_m * is * is * k0;
    } else {
        k0 = g * is; k1 = beta ? __ldg(beta + c) : 0.f; k2 = 0.f;
    }
    const size_t base = (size_t)plane * S;
    const long long s0 = (long long)blockIdx.x * kBnChunk, len = min((long long)kBnChunk, S - s0);
    if (((base + s0) | (size_t)len) % 4 == 0) {
        for (long long i = 4 * threadIdx.x; i < len; i += 4 * kBnThreads) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(x + base + s0 + i));
            float4 o;
            if (BWD) {
                const float4 d = __ldg(reinterpret_cast<const float4 *>(gy + base + s0 + i));
                o.x = d.x * k0 - k1 - (v.x - mu) * k2; o.y = d.y * k0 - k1 - (v.y - mu) * k2;
                o.z = d.z * k0 - k1 - (v.z - mu) * k2; o.w = d.w * k0 - k1 - (v.w - mu) * k2;
            } else {
                o.x = (v.x - mu) * k0 + k1; o.y = (v.y - mu) * k0 + k1; o.z = (v.z - mu) * k0 + k1; o.w = (v.w - mu) * k0 + k1;
            }
            *reinterpret_cast<float4 *>(out + base + s0 + i) = o;
        }
    } else {
        for (long long i = threadIdx.x; i < len; i += kBnThreads) {
            const float v = __ldg(x + base + s0 + i);
            out[base + s0 + i] = BWD ? __ldg(gy + base + s0 + i) * k0 - k1 - (v - mu) * k2 : (v - mu) * k0 + k1;
        }
    }
}

static int bn_splits(int N, int C, long long S)
{
    const long long chunks = ((S + kBnChunk - 1) / kBnChunk) * N;
    return (int)std::max<long long>(1, std::min<long long>(chunks, (148 * 8 + C - 1) / C));
}

}  // namespace side

using namespace side;

extern "C" size_t side_bn_train_ws_bytes(int N, int C, long long S)
{
    if (N <= 0 || C <= 0 || S <= 0) return 0;
    return sizeof(double) * 2 * (size_t)C * bn_splits(N, C, S) + sizeof(float) * 2 * (size_t)C;
}

extern "C" int side_bn_train_fwd(const float *x, const float *gamma, const float *beta, float *running_mean, float *running_var,
                                 float *y, float *save_mean, float *save_invstd, int N, int C, long long S, float eps, float momentum,
                                 void *ws, size_t ws_bytes, void *stream)
{
    SIDE_REQUIRE(N > 0 && C > 0 && S > 0 && (long long)N * C <= 0x7fffffffll && C <= 65535, "side_bn_train_fwd: bad shape");
    SIDE_REQUIRE((long long)N * C < 65536ll * 32768ll, "side_bn_train_fwd: too many planes");
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(y); SIDE_REQUIRE_DEV(save_mean); SIDE_REQUIRE_DEV(save_invstd);
    if (ws == nullptr || ws_bytes < side_bn_train_ws_bytes(N, C, S) || !is_device_ptr(ws) || (reinterpret_cast<uintptr_t>(ws) & 7)) {
        set_error("side_bn_train_fwd: needs side_bn_train_ws_bytes(...) bytes of 8-byte aligned device workspace");
        return SIDE_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int splits = bn_splits(N, C, S);
    double *partial = reinterpret_cast<double *>(ws);
    bn_stats_kernel<false><<<dim3(C, splits), kBnThreads, 0, st>>>(x, nullptr, nullptr, partial, N, C, S, splits);
    SIDE_LAUNCH_CHECK("bn_stats_kernel");
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(partial, C, splits, (double)N * (double)S, eps, momentum, save_mean,
                                                        save_invstd, running_mean, running_var);
    SIDE_LAUNCH_CHECK("bn_finalize_kernel");
    const long long planes = (long long)N * C;
    SIDE_REQUIRE(planes <= 65535, "side_bn_train_fwd: N * C must not exceed 65535");
    bn_apply_kernel<false><<<dim3((unsigned)ceil_div(S, kBnChunk), (unsigned)planes), kBnThreads, 0, st>>>(
        x, nullptr, gamma, beta, save_mean, save_invstd, nullptr, 0.f, y, C, S);
    SIDE_LAUNCH_CHECK("bn_apply_kernel");
    return SIDE_OK;
}

extern "C" int side_bn_train_bwd(const float *x, const float *gy, const float *gamma, const float *save_mean, const float *save_invstd,
                                 float *gx, float *ggamma, float *gbeta, int N, int C, long long S, void *ws, size_t ws_bytes,
                                 void *stream)
{
    SIDE_REQUIRE(N > 0 && C > 0 && S > 0 && C <= 65535 && (long long)N * C <= 65535, "side_bn_train_bwd: bad shape");
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(gy); SIDE_REQUIRE_DEV(save_mean); SIDE_REQUIRE_DEV(save_invstd);
    if (ws == nullptr || ws_bytes < side_bn_train_ws_bytes(N, C, S) || !is_device_ptr(ws) || (reinterpret_cast<uintptr_t>(ws) & 7)) {
        set_error("side_bn_train_bwd: needs side_bn_train_ws_bytes(...) bytes of 8-byte aligned device workspace");
        return SIDE_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int splits = bn_splits(N, C, S);
    double *partial = reinterpret_cast<double *>(ws);
    float *sums = reinterpret_cast<float *>(partial + 2 * (size_t)C * splits);
    bn_stats_kernel<true><<<dim3(C, splits), kBnThreads, 0, st>>>(x, gy, save_mean, partial, N, C, S, splits);
    SIDE_LAUNCH_CHECK("bn_bwd_stats_kernel");
    bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(partial, C, splits, save_invstd, sums, ggamma, gbeta);
    SIDE_LAUNCH_CHECK("bn_bwd_finalize_kernel");
    if (gx) {
        SIDE_REQUIRE_DEV(gx);
        bn_apply_kernel<true><<<dim3((unsigned)ceil_div(S, kBnChunk), (unsigned)((long long)N * C)), kBnThreads, 0, st>>>(
            x, gy, gamma, nullptr, save_mean, save_invstd, sums, (float)(1.0 / ((double)N * (double)S)), gx, C, S);
        SIDE_LAUNCH_CHECK("bn_bwd_apply_kernel");
    }
    return SIDE_OK;
}
