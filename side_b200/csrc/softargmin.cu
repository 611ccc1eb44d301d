// softargmin.cu -- fused AvgPool2d(S) + softmax over D + expectation (SURVEY.md section 8 row A7 tail).
// Replaces stereo_network_old.py:228-236 (AvgPool2d, softmax, then a Python loop of D multiply-adds).
// One warp per RoI: the S*S pooling is a per-lane serial sum, max / sum / expectation are warp-shuffle
// reductions, so the whole tail is one launch and each logit is read exactly once.
#include "common.cuh"

namespace side {

constexpr int kSamWarps = 4;

// Each lane owns candidates i = lane, lane+32, ... (D <= 32*kMaxPerLane).
constexpr int kMaxPerLane = 8;

__global__ void __launch_bounds__(kSamWarps * 32) softargmin_fwd_kernel(const float *__restrict__ logits,
                                                                       const float *__restrict__ depth_bin,
                                                                       float *__restrict__ depth,
                                                                       float *__restrict__ prob, int N, int D, int SS)
{
    const int n = blockIdx.x * kSamWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    const float *lg = logits + (size_t)n * D * SS;
    const float inv = 1.0f / (float)SS;
    float v[kMaxPerLane];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int i = lane + 32 * k;
        if (i < D) {
            float s = 0.f;
            for (int q = 0; q < SS; ++q) s += lg[(size_t)i * SS + q];
            v[k] = s * inv;
            mx = fmaxf(mx, v[k]);
        }
    }
    mx = warp_max(mx);
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int i = lane + 32 * k;
        if (i < D) {
            v[k] = expf(v[k] - mx);
            den += v[k];
        }
    }
    den = warp_sum(den);
    const float rden = 1.0f / den;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
        const int i = lane + 32 * k;
        if (i < D) {
            const float pv = v[k] * rden;
            if (prob) prob[(size_t)n * D + i] = pv;
            acc = fmaf(pv, depth_bin[(size_t)n * D + i], acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) depth[n] = acc;
}

// d depth / d logit_pool[i] = p_i (bin_i - depth);  each of the S*S pooled inputs gets 1/(S*S) of it.
__global__ void softargmin_bwd_kernel(const float *__restrict__ prob, const float *__restrict__ depth_bin,
                                      const float *__restrict__ depth, const float *__restrict__ gdepth,
                                      float *__restrict__ glogits, float *__restrict__ gdepth_bin, int N, int D, int SS)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)N * D) return;
    const int n = (int)(t / D);
    const float g = gdepth[n], pv = prob[t];
    const float gl = g * pv * (depth_bin[t] - depth[n]) / (float)SS;
    if (glogits)
        for (int q = 0; q < SS; ++q) glogits[t * SS + q] = gl;
    if (gdepth_bin) gdepth_bin[t] = g * pv;
}

}  // namespace side

using namespace side;

extern "C" int side_softargmin_fwd(const float *logits, const float *depth_bin, float *depth, float *prob, int N,
                                   int D, int S, void *stream)
{
    SIDE_REQUIRE(N >= 0 && D >= 1 && S >= 1, "side_softargmin_fwd: bad shape");
    SIDE_REQUIRE(D <= 32 * kMaxPerLane, "side_softargmin_fwd: D=%d exceeds %d", D, 32 * kMaxPerLane);
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(logits); SIDE_REQUIRE_DEV(depth_bin); SIDE_REQUIRE_DEV(depth);
    softargmin_fwd_kernel<<<ceil_div(N, kSamWarps), kSamWarps * 32, 0, (cudaStream_t)stream>>>(logits, depth_bin, depth,
                                                                                             prob, N, D, S * S);
    SIDE_LAUNCH_CHECK("softargmin_fwd_kernel");
    return SIDE_OK;
}

extern "C" int side_softargmin_bwd(const float *prob, const float *depth_bin, const float *depth, const float *gdepth,
                                   float *glogits, float *gdepth_bin, int N, int D, int S, void *stream)
{
    SIDE_REQUIRE(N >= 0 && D >= 1 && S >= 1, "side_softargmin_bwd: bad shape");
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(prob); SIDE_REQUIRE_DEV(depth_bin); SIDE_REQUIRE_DEV(depth); SIDE_REQUIRE_DEV(gdepth);
    softargmin_bwd_kernel<<<ceil_div((long long)N * D, 256), 256, 0, (cudaStream_t)stream>>>(
        prob, depth_bin, depth, gdepth, glogits, gdepth_bin, N, D, S * S);
    SIDE_LAUNCH_CHECK("softargmin_bwd_kernel");
    return SIDE_OK;
}
