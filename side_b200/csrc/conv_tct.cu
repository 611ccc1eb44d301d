// conv_tct.cu -- role-swapped variant of conv_tc.cu for Cout == 64 layers on 16x16 maps (dres0.0, dres0.3, dres1.0 of the
// aggregation network: 58 % of its FLOPs).
//
// Measured on B200 (tools/dbg_conv2.py, MMA-only mode): one tcgen05.mma.kind::tf32 with M = 128, K = 8 takes ~115 cycles
// whatever N is (16 ... 256); the 128 x N x 8 MACs only take 128 * N / 256 cycles.  With the voxels on M and Cout = 64 on N,
// conv_tc.cu issues MMAs of N = 128 (hi*hi | hi*lo) and N = 64 (lo*hi): 96 cycles of math in ~230.  Here the operands
// swap roles so that every instruction carries N = 256:
//     A (M = 128 rows)  = [W_hi ; W_lo]   64 + 64 output channels, K-major -- the weight tiles conv_tc_prepare already lays
//                          out back to back;   second instruction: [0 ; W_hi]  (a zero block kept in front of W_hi)
//     B (N = 256 rows)  = one full 16x16 slice of voxels of X_hi, then of X_lo
//     D [128 lanes x 256 columns]: lanes 0-63 = W_hi*X_hi (main), lanes 64-127 = W_lo*X_hi + W_hi*X_lo (cross terms, kept
//                          out of the main accumulator because the tensor core truncates on accumulation)
// 2 instructions of 128 cycles per k-step for 256 voxels instead of 2 x 115 for 128 voxels.
// The X operand uses the kh-view trick of conv_tc.cu: one TMA box with a one-row halo above and below (18 x 16 pixels)
// serves the three vertical taps (tap kh starts kh rows = kh * 2048 bytes further down, swizzle phase unchanged), so the
// activations cross L2 -> shared memory once per (kd, kw, channel block) instead of once per tap.
// Epilogue: TMEM lane = channel, column = voxel; a main warp and its cross warp exchange 32-column chunks through shared
// memory (alternating roles), add, apply the folded BatchNorm / ReLU / residual and store -- lane = channel means every
// store instruction writes 128 contiguous bytes of one voxel's channels-last row.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>

#include "tc_common.cuh"

namespace side {

constexpr int kTtThreads = 192;
constexpr int kTtCout = 64;
constexpr int kTtVox = 256;                         // voxels per tile = MMA N
constexpr uint32_t kTtWPart = kTtCout * 128;        // 8 KB: 64 rows x 32 tf32
constexpr uint32_t kTtWSlot = 3 * kTtWPart;         // [zeros][W_hi][W_lo]
constexpr int kTtWSlots = 2, kTtXSlots = 2;

struct ConvTtParams {
    const float *wp;                  // [nkb][2][64 x 32] swizzled weight tiles (hi, lo) from side_conv_tc_prep_weights
    float *y, *y_hi, *y_lo;
    const float *scale, *shift, *residual;
    int relu;
    int ncb, kd, ntiles, D, bw, bh;   // bw * bh == 256, one tile = one (sample, depth slice)
    int klast;                        // K-steps with data in the last channel block (2 when Cin % 64 == 32 on fp16, else 4)
    uint32_t x_part;                  // bytes of one half (hi or lo) of an X slot: (bh + 2) * bw * 128
    uint32_t *rs;                     // fp16 range-guard slot of this launch (tc_common.cuh) or NULL
    float acc_fix;                    // 1 + (MMA steps per accumulator) * 2^-26 (tc_common.cuh: truncating accumulation)
};

__device__ __forceinline__ void tt_tma_load_5d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, int c4,
                                               uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tt_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <bool F16>
__global__ void __launch_bounds__(kTtThreads, 1) conv_tct_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                                 const __grid_constant__ CUtensorMap tm_lo, ConvTtParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t fullX[kTtXSlots], emptyX[kTtXSlots], fullW[kTtWSlots], emptyW[kTtWSlots];
    __shared__ __align__(8) uint64_t tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float s_scale[kTtCout], s_shift[kTtCout];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler as well
    unsigned char *base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char *xring = base;                                               // kTtXSlots x [X_hi halo box][X_lo halo box]
    unsigned char *wring = xring + (size_t)kTtXSlots * 2 * p.x_part;           // kTtWSlots x [zeros][W_hi][W_lo]
    float *exch = reinterpret_cast<float *>(wring + (size_t)kTtWSlots * kTtWSlot);   // [2 pairs][2 buffers][32 cols][32 lanes]

    if (tid < kTtCout) {
        s_scale[tid] = p.scale ? p.scale[tid] : 1.0f;
        s_shift[tid] = p.shift ? p.shift[tid] : 0.0f;
    }
    for (int i = tid; i < kTtWSlots * (int)(kTtWPart / 16); i += kTtThreads) {      // the zero block in front of every W_hi
        const int s = i / (int)(kTtWPart / 16), o = i - s * (int)(kTtWPart / 16);
        *reinterpret_cast<float4 *>(wring + (size_t)s * kTtWSlot + (size_t)o * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&fullX[i], 1); mbar_init(&emptyX[i], 1);
            mbar_init(&fullW[i], 1); mbar_init(&emptyW[i], 1);
            mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();            // the zero blocks are read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const int pd = (p.kd - 1) / 2;

    if (warp == 0) {
        // ================= TMA producer =================
        // all 32 lanes run the loops (uniform control flow and operands), the elected lane issues -- see tc_elect_one()
        const bool leader = tc_elect_one();
        if (leader) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_lo) : "memory");
        }
        int sx = 0, sw = 0;
        uint32_t phx = 0, phw = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const int n = tile / p.D, d0 = tile - n * p.D;
            for (int kdi = 0; kdi < p.kd; ++kdi)
                for (int kwi = 0; kwi < 3; ++kwi)
                    for (int cb = 0; cb < p.ncb; ++cb) {
                        mbar_wait(&emptyX[sx], phx ^ 1u);
                        unsigned char *xs = xring + (size_t)sx * 2 * p.x_part;
                        if (leader) {
                            mbar_expect_tx(&fullX[sx], 2 * p.x_part);
                            tt_tma_load_5d(xs, &tm_hi, cb * (F16 ? 64 : 32), kwi - 1, -1, d0 + kdi - pd, n, &fullX[sx]);
                            tt_tma_load_5d(xs + p.x_part, &tm_lo, cb * (F16 ? 64 : 32), kwi - 1, -1, d0 + kdi - pd, n, &fullX[sx]);
                        }
                        if (++sx == kTtXSlots) { sx = 0; phx ^= 1u; }
                        for (int khi = 0; khi < 3; ++khi) {
                            const int kb = ((kdi * 3 + khi) * 3 + kwi) * p.ncb + cb;
                            mbar_wait(&emptyW[sw], phw ^ 1u);
                            if (leader) {
                                mbar_expect_tx(&fullW[sw], 2 * kTtWPart);
                                bulk_g2s(wring + (size_t)sw * kTtWSlot + kTtWPart, p.wp + (size_t)kb * (2 * kTtWPart / 4), 2 * kTtWPart,
                                         &fullW[sw]);
                            }
                            if (++sw == kTtWSlots) { sw = 0; phw ^= 1u; }
                        }
                    }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // all 32 lanes run the loops, the elected lane issues the MMAs and commits (tc_elect_one(): descriptors stay in uniform
        // registers, no per-instruction election loop)
        const bool leader = tc_elect_one();
        const uint32_t idesc = tc_idesc<F16>(128, kTtVox);
        const uint32_t view = (uint32_t)p.bw * 128u;              // bytes per halo row
        const int ngroups = p.kd * 3 * p.ncb;
        const uint64_t xdesc0 = tc_smem_desc(smem_u32(xring)), wdesc0 = tc_smem_desc(smem_u32(wring));
        int sx = 0, sw = 0, acc = 0, cbm = 0;
        uint32_t phx = 0, phw = 0, acc_ph = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[acc], acc_ph ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kTtVox);
            for (int g = 0; g < ngroups; ++g) {
                mbar_wait(&fullX[sx], phx);
                tc_fence_after();
                const uint64_t xs = tc_desc_add(xdesc0, (uint32_t)sx * 2u * p.x_part);
                const int ksteps = (cbm == p.ncb - 1) ? p.klast : 4;
                if (++cbm == p.ncb) cbm = 0;
                for (int khi = 0; khi < 3; ++khi) {
                    mbar_wait(&fullW[sw], phw);
                    tc_fence_after();
                    const uint64_t a_zh0 = tc_desc_add(wdesc0, (uint32_t)sw * kTtWSlot);       // [0 ; W_hi]
                    const uint64_t a_hl0 = tc_desc_add(a_zh0, kTtWPart);                       // [W_hi ; W_lo]
                    const uint64_t b_hi0 = tc_desc_add(xs, (uint32_t)khi * view), b_lo0 = tc_desc_add(b_hi0, p.x_part);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k >= ksteps) break;
                            tc_mma<F16>(tmem_d, tc_desc_add(a_hl0, k * 32), tc_desc_add(b_hi0, k * 32), idesc, (g | khi | k) != 0 ? 1u : 0u);
                            tc_mma<F16>(tmem_d, tc_desc_add(a_zh0, k * 32), tc_desc_add(b_lo0, k * 32), idesc, 1u);
                        }
                        tc_commit(&emptyW[sw]);
                    }
                    if (++sw == kTtWSlots) { sw = 0; phw ^= 1u; }
                }
                if (leader) tc_commit(&emptyX[sx]);
                if (++sx == kTtXSlots) { sx = 0; phx ^= 1u; }
            }
            if (leader) tc_commit(&tmem_full[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
        __syncwarp();
    } else {
        // ================= epilogue =================
        const int lg = warp & 3;                 // TMEM lanes 32 lg .. 32 lg + 31: lg 0,1 = main ch 0-31 / 32-63; lg 2,3 = cross
        const int pair = lg & 1;                 // (lg, lg ^ 2) hold the two parts of the same 32 channels
        const bool is_main = lg < 2;
        const int ch = 32 * pair + lane;
        float *ex = exch + (size_t)pair * 2 * 1024;
        const float sc = s_scale[ch], sh = s_shift[ch];
        int acc = 0;
        uint32_t acc_ph = 0;
        float amax = 0.f;                        // range guard: max |x| of what this thread splits into fp16 pairs
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            mbar_wait(&tmem_full[acc], acc_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * kTtVox);
            const size_t vox0 = (size_t)tile * kTtVox;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                float v[32];
                tt_ld32(taddr + (uint32_t)(c * 32), v);
                if (is_main) {                                        // undo the accumulator's truncation bias (main terms only)
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= p.acc_fix;
                } else if (F16) {                                     // cross terms carry the 2^11 of the scaled lo operands
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= kF16LoInv;
                }
                float *eb = ex + (c & 1) * 1024;
                const bool store_role = is_main == ((c & 1) == 0);    // even chunks: the main warp finishes, odd: the cross warp
                if (!store_role) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) eb[j * 32 + lane] = v[j];
                }
                if (pair == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
                else asm volatile("bar.sync 2, 64;" ::: "memory");
                if (store_role) {
                    const size_t o0 = (vox0 + (size_t)c * 32) * kTtCout + ch;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float o = fmaf(v[j] + eb[j * 32 + lane], sc, sh);
                        if (p.relu == 1) o = fmaxf(o, 0.f);
                        if (p.residual) o += __ldg(p.residual + o0 + (size_t)j * kTtCout);
                        if (p.relu == 2) o = fmaxf(o, 0.f);
                        if (p.y) p.y[o0 + (size_t)j * kTtCout] = o;
                        if (F16 && p.y_hi) {
                            __half h, l;
                            amax = fmaxf(amax, fabsf(o));
                            f16_split(o, h, l);
                            reinterpret_cast<__half *>(p.y_hi)[o0 + (size_t)j * kTtCout] = h;
                            reinterpret_cast<__half *>(p.y_lo)[o0 + (size_t)j * kTtCout] = l;
                        } else if (p.y_hi) {
                            const float h = tf32_hi(o);
                            p.y_hi[o0 + (size_t)j * kTtCout] = h;
                            p.y_lo[o0 + (size_t)j * kTtCout] = o - h;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
        if (F16 && p.rs && p.y_hi) range_commit(p.rs, amax);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// shared with conv_tc.cu
int conv_make_act_tmap(CUtensorMap *tm, const float *base, int Nn, int D, int H, int W, int C, int bd, int bh, int bw, int sh,
                       int sw, int halo_h, bool f16);

bool conv_tct_supported(int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int stride, bool f16)
{
    if (Cout != kTtCout || Cin % 32 || stride != 1 || kh != 3 || kw != 3 || (kd != 1 && kd != 3)) return false;
    if (H * W != kTtVox || (W % 8) || W > 64) return false;          // one tile = one full depth slice
    const size_t smem = (size_t)kTtXSlots * 2 * (size_t)(H + 2) * W * 128 + (size_t)kTtWSlots * kTtWSlot + 2 * 2 * 1024 * 4 + 1024;
    return smem <= 212 * 1024;
}

int conv_tct_launch(const float *x_hi, const float *x_lo, const float *wp, const float *scale, const float *shift,
                    const float *residual, float *y, float *y_hi, float *y_lo, int Nn, int D, int H, int W, int Cin, int kd,
                    int relu, int sm_count, cudaStream_t st, bool f16)
{
    CUtensorMap tm_hi, tm_lo;
    int rc;
    if ((rc = conv_make_act_tmap(&tm_hi, x_hi, Nn, D, H, W, Cin, 1, H, W, 1, 1, 1, f16))) return rc;
    if ((rc = conv_make_act_tmap(&tm_lo, x_lo, Nn, D, H, W, Cin, 1, H, W, 1, 1, 1, f16))) return rc;
    ConvTtParams p;
    p.wp = wp; p.y = y; p.y_hi = y_hi; p.y_lo = y_lo; p.scale = scale; p.shift = shift; p.residual = residual; p.relu = relu;
    p.ncb = f16 ? (Cin + 63) / 64 : Cin / 32; p.kd = kd; p.klast = (f16 && Cin % 64 == 32) ? 2 : 4; p.ntiles = Nn * D; p.D = D; p.bw = W; p.bh = H;
    p.x_part = (uint32_t)(H + 2) * W * 128u;
    p.rs = f16 ? range_slot_next() : nullptr;
    p.acc_fix = tc_acc_fix(kd * 9 * ((p.ncb - 1) * 4 + p.klast));
    const size_t smem = (size_t)kTtXSlots * 2 * p.x_part + (size_t)kTtWSlots * kTtWSlot + 2 * 2 * 1024 * 4 + 1024;
    if ((rc = set_smem_attr(f16 ? (const void *)conv_tct_kernel<true> : (const void *)conv_tct_kernel<false>, smem))) return rc;
    const unsigned grid = (unsigned)std::min(p.ntiles, sm_count);
    if (f16) conv_tct_kernel<true><<<grid, kTtThreads, smem, st>>>(tm_hi, tm_lo, p);
    else conv_tct_kernel<false><<<grid, kTtThreads, smem, st>>>(tm_hi, tm_lo, p);
    SIDE_LAUNCH_CHECK("conv_tct_kernel");
    return SIDE_OK;
}

}  // namespace side
