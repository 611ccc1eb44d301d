// common.cuh -- shared helpers for libside_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/side_b200.h"

namespace side {

void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

inline int cuda_fail(cudaError_t e, const char *what)
{
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SIDE_ERR_CUDA;
}

// NCHW -> NHWC staging copy (layout.cu)
int launch_nchw_to_nhwc(const float *in, float *out, int B, int C, int HW, cudaStream_t st);
int launch_nhwc_to_nchw(const float *in, float *out, int B, int C, int HW, cudaStream_t st);

// true when p is device (or managed) memory of the current context; NULL is handled by callers
bool is_device_ptr(const void *p);

#define SIDE_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            side::set_error(__VA_ARGS__);       \
            return SIDE_ERR_INVALID_ARG;        \
        }                                       \
    } while (0)

#define SIDE_REQUIRE_DEV(p)                                              \
    do {                                                                 \
        if ((p) == nullptr) {                                            \
            side::set_error("%s: null pointer", #p);                     \
            return SIDE_ERR_INVALID_ARG;                                 \
        }                                                                \
        if (!side::is_device_ptr(p)) {                                   \
            side::set_error("%s: not a device pointer (no CPU path)", #p); \
            return SIDE_ERR_NOT_DEVICE;                                  \
        }                                                                \
    } while (0)

#define SIDE_LAUNCH_CHECK(name)                                       \
    do {                                                              \
        side::g_launches.fetch_add(1, std::memory_order_relaxed);     \
        cudaError_t e__ = cudaGetLastError();                         \
        if (e__ != cudaSuccess) return side::cuda_fail(e__, name);    \
    } while (0)

#define SIDE_CUDA(call)                                               \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return side::cuda_fail(e__, #call);   \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// opt a kernel into > 48 KB of dynamic shared memory (idempotent, thread-safe in the runtime)
inline int set_smem_attr(const void *func, size_t bytes)
{
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    }
    return SIDE_OK;
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming (evict-first) stores for write-once volumes: keeps the L2 for the feature maps
__device__ __forceinline__ void st_cs(float4 *p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(float *p, float v) { __stcs(p, v); }

// sigmoid with the same formula torch uses (1 / (1 + exp(-x))), accurate expf
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------------
// mbarrier + TMA 1-D bulk copy (cp.async.bulk -> SASS UBLKCP).  Waits are BOUNDED: a barrier that never
// completes traps (CUDA error) instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // try_wait suspends for a HW-defined time slice per call; ~2^22 tries is seconds, far beyond any legal wait
    for (uint32_t it = 0; it < (1u << 22); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace side
