// tc_common.cuh -- tcgen05 / TMEM helpers shared by the tensor-core kernels (dcn_fwd_tc.cu, conv_tc.cu).
#pragma once
#include <cuda_fp16.h>

#include "dcn_common.cuh"

namespace side {

constexpr int kTcBK = 32;             // tf32 elements per stage row (= 128 bytes, one swizzle row)

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// K-major, SWIZZLE_128B operand tile: 8-row groups are 1024 bytes apart (SBO), LBO is unused for swizzled K-major
// (encoded as 1 like CUTLASS does), descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}

// kind::f16 (fp16 operands, fp32 accumulate): K = 16 per instruction = the same 32 bytes per row as kind::tf32's K = 8,
// so tiles, swizzle and descriptor stepping are identical; it runs at twice the tf32 rate.
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
template <bool F16>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    if (F16) tc_mma_f16(tmem_d, adesc, bdesc, idesc, accum);
    else tc_mma_tf32(tmem_d, adesc, bdesc, idesc, accum);
}

// One lane of a converged warp, the same one every time for a full mask.  The MMA-issuing warps run their loops with ALL lanes
// (warp-uniform control flow and operands) and only predicate the tcgen05 instructions on this: with the whole loop inside
// `if (lane == 0)` the compiler cannot prove the descriptors uniform, moves each of them through R2UR and wraps every
// tcgen05.mma in an ELECT / BRA.U.ANY loop -- ~45 instructions of one thread per K-step, which (not the tensor core) was the
// "~115 cycles per MMA whatever N" floor measured in round 1.
__device__ __forceinline__ bool tc_elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// shared-memory descriptor advanced by `bytes` (a multiple of 16 that stays inside the tile: the 14-bit address field cannot carry)
__device__ __forceinline__ uint64_t tc_desc_add(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// "3xFP16" operand split, the fp16 counterpart of the tf32 hi/lo split: x = hi + lo' * 2^-11 with hi = fp16(x) and
// lo' = fp16((x - hi) * 2^11) -- 11 + 11 significand bits like two tf32 values; the 2^11 keeps lo' out of the fp16
// subnormals.  hi*hi goes to the main accumulator, hi*lo' + lo'*hi to the cross accumulator, which the epilogue scales by
// 2^-11.  Saturating conversions: |x| must stay below 65504 (activations behind a BatchNorm do).
constexpr float kF16LoScale = 2048.0f, kF16LoInv = 1.0f / 2048.0f;
__device__ __forceinline__ __half f16_sat(float v)
{
    unsigned short r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
    return __ushort_as_half(r);
}
__device__ __forceinline__ void f16_split(float v, __half &h, __half &l)
{
    h = f16_sat(v);
    l = f16_sat((v - __half2float(h)) * kF16LoScale);
}

// ------------------------------------------------------------------------------------------------
// Range guard of the fp16 pairs.  fp16 resolves 2^-24 .. 65504 where the reference's fp32 covers 2^-149 .. 3.4e38: a tensor
// whose largest element reaches 65504 saturates, one whose largest element is below 2^-18 loses more than the 1e-4 parity bar
// to the subnormal spacing of `lo'` (absolute error 2^-36).  Every kernel that WRITES fp16 pairs tracks max |x| of what it
// splits and folds it into one 32-bit slot of a device status ring registered with side_tc_range_guard(): the host hands
// each guarded launch the next slot (round robin), a warp only issues its atomicMax when its maximum beats what the slot
// already holds (a handful of atomics per launch, no ordering between launches needed).  The host classifies the slots when
// it wants to know (side_b200.ops.tc_range_status: any slot >= 65504 -> saturated, any non-zero slot < 2^-18 -> underflow),
// zeroes them and, if a flag is up, repeats the work in 3xTF32.
// ------------------------------------------------------------------------------------------------
uint32_t *range_slot_next();                       // api.cu: next slot of the ring registered for the current device, or nullptr

// all 32 lanes of a warp call this once at the end of the kernel with the maximum |x| they split
__device__ __forceinline__ void range_commit(uint32_t *slot, float amax)
{
    amax = warp_max(amax);
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0) {
        const uint32_t b = __float_as_uint(amax);  // non-negative floats order like their bit patterns; +inf counts as saturation
        uint32_t cur;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(slot) : "memory");
        if (b > cur) atomicMax(slot, b);
    }
}

// split two values and pack the halves pairwise (little endian: a in the low 16 bits).  Packed conversions: one
// cvt.rn.satfinite.f16x2.f32 per pair of values for hi and for lo (same rounding as the scalar form, 8 instead of 12 instructions)
__device__ __forceinline__ uint32_t f16x2_sat(float a, float b)
{
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void f16_split2(float a, float b, uint32_t &hi, uint32_t &lo)
{
    hi = f16x2_sat(a, b);
    const __half2 h2 = *reinterpret_cast<const __half2 *>(&hi);
    lo = f16x2_sat((a - __low2float(h2)) * kF16LoScale, (b - __high2float(h2)) * kF16LoScale);
}

__device__ __forceinline__ void tc_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

// The tensor core adds every MMA's partial sum into its fp32 accumulator with truncation toward zero.  Measured on B200
// (tools/conv_tc_err.py, 14 layer shapes, kind::tf32 and kind::f16 alike): outputs shrink by 1.35e-8 .. 1.9e-8 = 2^-26 of their
// magnitude per MMA step issued into the accumulator -- a systematic, data-independent bias (-9e-6 for a 512-channel 3x3
// layer) that adds up coherently over the ~130 layers of the network, while fp32 FMA rounding errors average out.  The
// epilogues multiply the main accumulator by 1 + steps * 2^-26 to remove the expected shrink (the small cross-term
// accumulators do not need it: they are 2^-11 of the result).
constexpr float kAccTruncPerStep = 1.4901161e-8f;     // 2^-26
__host__ __device__ __forceinline__ float tc_acc_fix(int mma_steps) { return 1.0f + (float)mma_steps * kAccTruncPerStep; }

// byte offset of (row, 16-byte chunk) inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128(int row, int chunk)
{
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// 16 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// instruction descriptor, kind::tf32: D = fp32, A = B = tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ __forceinline__ uint32_t tc_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16: A = B = fp16 (format 0), D = fp32
__host__ __device__ __forceinline__ uint32_t tc_idesc_f16(int M, int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <bool F16>
__host__ __device__ __forceinline__ uint32_t tc_idesc(int M, int N)
{
    return F16 ? tc_idesc_f16(M, N) : tc_idesc_tf32(M, N);
}

// w [Cout, Cin, KK] -> per-K-block tiles wp[kb][part][Cout x 32] in the swizzled shared-memory image (dcn_fwd_tc.cu)
int launch_tc_weight_prep(const float *w, float *wp, int Cout, int Cin, int KK, int split, cudaStream_t st);

}  // namespace side
