// layout.cu -- NCHW -> NHWC staging copy shared by the gather kernels (instance cost volume, tcgen05 DCN).
// Bilinear gathers fetch one pixel's channels at a time: channels-last makes every tap a contiguous 16-byte
// (4-channel) load and every warp request a set of full 128-byte lines.
#include <algorithm>
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

// channel concatenation of channels-last tensors [rows, C_i] -> [rows, sum C_i] in 16-byte pieces (torch.cat along the last
// dimension runs an element-wise strided copy per source at a third of the copy bandwidth)
struct CatParams {
    const uint4 *src[8];
    int c16[8];                 // 16-byte pieces per row of each source
    int off16[8];               // first piece of each source inside a destination row
    int nsrc, tot16;
};
__global__ void __launch_bounds__(256) cl_concat_kernel(CatParams p, uint4 *__restrict__ dst, long long rows)
{
    const long long total = rows * p.tot16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / p.tot16;
        const int c = (int)(i - r * p.tot16);
        int s = 0;
#pragma unroll
        for (int k = 1; k < 8; ++k)
            if (k < p.nsrc && c >= p.off16[k]) s = k;
        dst[i] = __ldg(p.src[s] + r * p.c16[s] + (c - p.off16[s]));
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

extern "C" int side_cl_concat(const void *const *srcs, const int *row_bytes, int nsrc, void *dst, long long rows, void *stream)
{
    using namespace side;
    SIDE_REQUIRE(nsrc >= 1 && nsrc <= 8 && rows >= 0 && srcs && row_bytes, "side_cl_concat: 1..8 sources");
    if (rows == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(dst);
    CatParams p{};
    int off = 0;
    for (int i = 0; i < nsrc; ++i) {
        SIDE_REQUIRE(row_bytes[i] > 0 && row_bytes[i] % 16 == 0, "side_cl_concat: rows must be multiples of 16 bytes");
        SIDE_REQUIRE_DEV(srcs[i]);
        p.src[i] = reinterpret_cast<const uint4 *>(srcs[i]);
        p.c16[i] = row_bytes[i] / 16;
        p.off16[i] = off;
        off += p.c16[i];
    }
    p.nsrc = nsrc; p.tot16 = off;
    const long long total = rows * off;
    cl_concat_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        p, reinterpret_cast<uint4 *>(dst), rows);
    SIDE_LAUNCH_CHECK("cl_concat_kernel");
    return SIDE_OK;
}
