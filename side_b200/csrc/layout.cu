// layout.cu -- NCHW -> NHWC staging copy shared by the gather kernels (instance cost volume, tcgen05 DCN).
// Bilinear gathers fetch one pixel's channels at a time: channels-last makes every tap a contiguous 16-byte
// (4-channel) load and every warp request a set of full 128-byte lines.
#include "common.cuh"

namespace side {

__global__ void nchw_to_nhwc_kernel(const float *__restrict__ in, float *__restrict__ out, int C, int HW)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *ip = in + (size_t)b * C * HW;
    float *op = out + (size_t)b * C * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? __ldg(ip + (size_t)c * HW + p) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        if (p < HW && c < C) op[(size_t)p * C + c] = tile[threadIdx.x][i];
    }
}


// channels-last [B, HW, C] -> NCHW [B, C, HW] (tensor-core convolution outputs handed back to NCHW consumers)
// ld: floats per channels-last row (>= C: the first C channels of every row are taken)
__global__ void nhwc_to_nchw_kernel(const float *__restrict__ in, float *__restrict__ out, int C, int HW, int ld)
{
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float *ip = in + (size_t)b * ld * HW;
    float *op = out + (size_t)b * C * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int p = p0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < HW && c < C) ? __ldg(ip + (size_t)p * ld + c) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, p = p0 + threadIdx.x;
        if (c < C && p < HW) op[(size_t)c * HW + p] = tile[threadIdx.x][i];
    }
}

int launch_nchw_to_nhwc(const float *in, float *out, int B, int C, int HW, cudaStream_t st)
{
    dim3 tg(ceil_div(HW, 32), ceil_div(C, 32), B), tb(32, 8);
    SIDE_REQUIRE(B <= 65535 && ceil_div(C, 32) <= 65535, "nchw_to_nhwc: grid too large");
    nchw_to_nhwc_kernel<<<tg, tb, 0, st>>>(in, out, C, HW);
    SIDE_LAUNCH_CHECK("nchw_to_nhwc_kernel");
    return SIDE_OK;
}

int launch_nhwc_to_nchw(const float *in, float *out, int B, int C, int HW, cudaStream_t st)
{
    dim3 tg(ceil_div(HW, 32), ceil_div(C, 32), B), tb(32, 8);
    SIDE_REQUIRE(B <= 65535 && ceil_div(C, 32) <= 65535, "nhwc_to_nchw: grid too large");
    nhwc_to_nchw_kernel<<<tg, tb, 0, st>>>(in, out, C, HW, C);
    SIDE_LAUNCH_CHECK("nhwc_to_nchw_kernel");
    return SIDE_OK;
}

}  // namespace side

extern "C" int side_cl_to_nchw_ld(const float *x, int ld, float *y, int B, int C, long long HW, void *stream);
extern "C" int side_cl_to_nchw(const float *x, float *y, int B, int C, long long HW, void *stream)
{
    return side_cl_to_nchw_ld(x, C, y, B, C, HW, stream);
}

extern "C" int side_cl_to_nchw_ld(const float *x, int ld, float *y, int B, int C, long long HW, void *stream)
{
    using namespace side;
    SIDE_REQUIRE(B >= 0 && C > 0 && ld >= C && HW > 0 && HW < (1ll << 31), "side_cl_to_nchw: bad shape");
    if (B == 0) return SIDE_OK;
    SIDE_REQUIRE(B <= 65535 && ceil_div(C, 32) <= 65535, "side_cl_to_nchw: grid too large");
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(y);
    dim3 tg(ceil_div(HW, 32), ceil_div(C, 32), B), tb(32, 8);
    nhwc_to_nchw_kernel<<<tg, tb, 0, (cudaStream_t)stream>>>(x, y, C, (int)HW, ld);
    SIDE_LAUNCH_CHECK("nhwc_to_nchw_kernel");
    return SIDE_OK;
}
