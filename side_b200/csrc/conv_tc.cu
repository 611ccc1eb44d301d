// conv_tc.cu -- 3x3x3 (and 3x3) stride-1 "same" convolutions of the instance-depth aggregation network on the
// 5th-gen tensor cores (cost_volume.dres0/dres1/dres2/classify,
// stereo_network_old.py:139-171, 205-227; nn.Conv3d(3, padding=1, bias=False) + BatchNorm3d + ReLU).
//
// cuDNN serves these fp32 convolutions with FFT / Winograd / SIMT-SGEMM kernels (about 70 % of an inference step,
// profiles/r1_launches_step_3xtf32.md).  Here each one is an implicit GEMM on tcgen05:
//     y[voxel, o] = sum_{tap, c} x[voxel + tap, c] * W[o, c, tap]        M = N*D*H*W voxels, N = Cout, K = 27*Cin
// with channels-last activations (NDHWC), so that the A operand of one (tap, 32-channel block) for a tile of 128
// voxels is ONE 5-D TMA box {32 ch, bw, bh, bd, 1} whose start coordinate is shifted by the tap: the hardware's
// out-of-bound zero fill IS the convolution's zero padding, and the 128-byte-swizzled box lands in shared memory in
// exactly the K-major SWIZZLE_128B layout tcgen05.mma reads.  No im2col, no gather warps.
//
//   warp 0   TMA producer : per k-block two tensor loads (hi and lo halves of the activations) + one bulk copy of the
//                           pre-swizzled weight tile, all completing on the stage's mbarrier
//   warp 1   MMA issuer   : one thread, tcgen05.mma.cta_group::1.kind::tf32, M=128, N=Cout, accumulators in TMEM,
//                           DOUBLE-BUFFERED (2 x 2 x Cout columns) so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2-5 epilogue    : tcgen05.ld -> folded BatchNorm -> ReLU -> (+ residual) -> fp32 channels-last stores, and/or
//                           the hi/lo tf32 split the next convolution consumes
// Persistent grid (one CTA per SM, static round-robin over tiles).
//
// Precision: 3xTF32 as in dcn_fwd_tc.cu -- x = hi + lo with hi = top 19 bits (exact split), products hi*hi + lo*hi +
// hi*lo accumulate in fp32; the dropped lo*lo term is 2^-22 relative.  Activations are kept pre-split in HBM
// (the producer layer's epilogue writes both halves), so no warp touches the operands between TMA and the MMA.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>

#include "tc_common.cuh"

namespace side {

constexpr int kCvThreads = 320;        // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane group, alternating 16-column chunks)
constexpr int kCvMaxStages = 6;
constexpr int kCvBM = 128;
constexpr int kCvMaxCout = 1536;       // all n-tiles' folded-BN vectors live in shared memory
constexpr uint32_t kCvATile = kCvBM * 128;   // 16 KB: 128 voxels x 32 tf32

struct ConvTcParams {
    const float *wp;                  // [nkb][2][Cout x 32] swizzled weight tiles (hi, lo)
    float *y, *y_hi, *y_lo;           // [M, Cout] channels-last outputs, any may be NULL
    const float *scale, *shift;       // folded BatchNorm, NULL = identity
    const float *residual;            // [M, Cout] added after the ReLU (dres2(cost) + cost), NULL = none
    int relu;
    int N, ncb, nkb, ntiles;          // N = output channels of ONE n-tile (<= 128); ntiles = m-tiles * n_ntiles
    int pool;                         // 1: MaxPool3d((1,2,2)) fused into the epilogue, outputs are [N, D, H/2, W/2, Cout]
    int klast;                        // MMA K-steps (of 32 bytes) that carry data in the LAST channel block (4, or 2 when Cin % 64 == 32 on fp16)
    int Ntot, n_ntiles;               // all output channels, number of n-tiles (Ntot = n_ntiles * N)
    long long ldy;                    // row stride of the outputs in floats (Ntot unless a column block of a wider matrix is written)
    int D, H, W;                      // spatial extent of one sample
    int bw, bh, bd;                   // TMA box extent along w, h, d; bd*bh*bw == 128
    int wt, ht, dt;                   // boxes per row (W / bw), per column (H / bh) and along depth (ceil(D / bd): the last box of a
                                      // sample may hang over the end -- TMA zero-fills the loads, the epilogue skips those voxels)
    int kd, kh, kw;                   // kernel extent (3,3,3), (1,3,3) or (1,1,1)
    int sh, sw;                       // stride along h / w (1 or 2); D, H, W above are OUTPUT extents
    int stages;
    // "kh view" mode (3x3 / 3x3x3 kernels, stride 1, bd == 1, bw % 8 == 0): one TMA box with a one-row halo above and below
    // ((bh + 2) x bw pixels) serves the three vertical taps -- tap kh reads the box from row kh on, a start address that
    // is a multiple of 1024 bytes, so the 128-byte swizzle phase is unchanged.  A ring (a_slots) and B ring (b_slots)
    // are then separate: A is loaded once per (kd, kw, channel block), B once per tap.
    int khv, a_slots, b_slots;
    uint32_t *rs;                     // fp16 range-guard slot of this launch (tc_common.cuh) or NULL
    float acc_fix;                    // 1 + (MMA steps per accumulator) * 2^-26: undoes the accumulator's truncation bias
    int dbg;                          // timing experiments only (side_conv_tc_set_mode bits 1, 2): skip the B / A copies
    uint32_t a_part;                  // bytes of one half (hi or lo) of an A slot
    int stg;                          // 1: epilogue staged through shared memory (coalesced global accesses), see the epilogue
    uint32_t stg_off;                 // byte offset of the staging area behind the operand ring
};
constexpr int kCvStgRowF = 36;                              // floats per staged row: 32 columns + 4 pad (conflict-free 16-byte accesses)
constexpr uint32_t kCvStgBytes = 8u * 32u * kCvStgRowF * 4u;   // 8 epilogue warps x 32 rows

// m-tile -> first voxel coordinates of its box
__device__ __forceinline__ void conv_tile_origin(const ConvTcParams &p, int mt, int &n, int &d0, int &h0, int &w0)
{
    const int tps = p.dt * p.ht * p.wt;                // tiles per sample
    n = mt / tps;
    int r = mt - n * tps;
    const int wi = r % p.wt;
    r /= p.wt;
    const int hi = r % p.ht;
    w0 = wi * p.bw; h0 = hi * p.bh; d0 = (r / p.ht) * p.bd;
}

__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

template <bool F16>
__global__ void __launch_bounds__(kCvThreads, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                                const __grid_constant__ CUtensorMap tm_lo, ConvTcParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[kCvMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kCvMaxStages];
    __shared__ __align__(8) uint64_t fullA[2], emptyA[2];       // kh-view mode: A ring (full_bar / empty_bar are the B ring)
    __shared__ __align__(8) uint64_t tmem_full[2];
    __shared__ __align__(8) uint64_t tmem_empty[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float s_scale[kCvMaxCout], s_shift[kCvMaxCout];

    constexpr int kKel = F16 ? 64 : 32;      // channels per 128-byte k-block row
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler as well
    const int N = p.N;
    const uint32_t b_part = (uint32_t)N * 128u;
    const uint32_t stage_bytes = 2 * kCvATile + 2 * b_part;
    unsigned char *tiles = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int stages = p.stages;

    for (int i = tid; i < p.Ntot; i += kCvThreads) {
        s_scale[i] = p.scale ? p.scale[i] : 1.0f;
        s_shift[i] = p.shift ? p.shift[i] : 0.0f;
    }
    if (tid == 0) {
        for (int i = 0; i < kCvMaxStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 8);
            mbar_init(&fullA[i], 1);
            mbar_init(&emptyA[i], 1);
        }
        mbar_fence_init();
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 4 * N) tmem_cols <<= 1;   // 2 buffers x (main, cross) accumulators
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int pd = (p.kd - 1) / 2, ph = (p.kh - 1) / 2, pw = (p.kw - 1) / 2;

    if (warp == 0) {
        // ================= TMA producer =================
        // All 32 lanes run the loops (uniform control flow, uniform operands); the elected lane issues -- see tc_elect_one().
        const bool leader = tc_elect_one();
        if (leader) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm_lo) : "memory");
        }
        int st = 0;
        uint32_t phs = 0;
        if (p.khv) {
            unsigned char *bring = tiles + (size_t)p.a_slots * 2 * p.a_part;
            int sa_i = 0;
            uint32_t pha = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int mt = tile / p.n_ntiles, nt = tile - mt * p.n_ntiles;
                int n, d0, h0, w0;
                conv_tile_origin(p, mt, n, d0, h0, w0);
                const float *wpt = p.wp + (size_t)nt * p.nkb * (2 * b_part / 4);
                for (int kdi = 0; kdi < p.kd; ++kdi)
                    for (int kwi = 0; kwi < 3; ++kwi)
                        for (int cb = 0; cb < p.ncb; ++cb) {
                            mbar_wait(&emptyA[sa_i], pha ^ 1u);
                            unsigned char *sa = tiles + (size_t)sa_i * 2 * p.a_part;
                            if (leader) {
                                mbar_expect_tx(&fullA[sa_i], 2 * p.a_part);
                                tma_load_5d(sa, &tm_hi, cb * kKel, w0 + kwi - 1, h0 - 1, d0 + kdi - pd, n, &fullA[sa_i]);
                                tma_load_5d(sa + p.a_part, &tm_lo, cb * kKel, w0 + kwi - 1, h0 - 1, d0 + kdi - pd, n, &fullA[sa_i]);
                            }
                            if (++sa_i == p.a_slots) { sa_i = 0; pha ^= 1u; }
                            for (int khi = 0; khi < 3; ++khi) {
                                const int kb = ((kdi * 3 + khi) * 3 + kwi) * p.ncb + cb;
                                mbar_wait(&empty_bar[st], phs ^ 1u);
                                if (leader) {
                                    mbar_expect_tx(&full_bar[st], 2 * b_part);
                                    bulk_g2s(bring + (size_t)st * 2 * b_part, wpt + (size_t)kb * (2 * b_part / 4), 2 * b_part, &full_bar[st]);
                                }
                                if (++st == p.b_slots) { st = 0; phs ^= 1u; }
                            }
                        }
            }
        } else {
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int mt = tile / p.n_ntiles, nt = tile - mt * p.n_ntiles;
                int n, d0, h0, w0;
                conv_tile_origin(p, mt, n, d0, h0, w0);
                const float *wpt = p.wp + (size_t)nt * p.nkb * (2 * b_part / 4);
                // (kd, kh, kw, channel block) carried incrementally: no integer divisions on the critical path of every k-block
                int kdi = 0, khi = 0, kwi = 0, cb = 0;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&empty_bar[st], phs ^ 1u);
                    unsigned char *sa = tiles + (size_t)st * stage_bytes;
                    const int cw = w0 * p.sw + kwi - pw, ch = h0 * p.sh + khi - ph;     // input coordinates of the box origin
                    if (leader) {
                        mbar_expect_tx(&full_bar[st], ((p.dbg & 4) ? 0u : 2 * kCvATile) + ((p.dbg & 2) ? 0u : 2 * b_part));
                        if (!(p.dbg & 4)) {
                            tma_load_5d(sa, &tm_hi, cb * kKel, cw, ch, d0 + kdi - pd, n, &full_bar[st]);
                            tma_load_5d(sa + kCvATile, &tm_lo, cb * kKel, cw, ch, d0 + kdi - pd, n, &full_bar[st]);
                        }
                        if (!(p.dbg & 2)) bulk_g2s(sa + 2 * kCvATile, wpt + (size_t)kb * (2 * b_part / 4), 2 * b_part, &full_bar[st]);
                    }
                    if (++st == stages) { st = 0; phs ^= 1u; }
                    if (++cb == p.ncb) {
                        cb = 0;
                        if (++kwi == p.kw) {
                            kwi = 0;
                            if (++khi == p.kh) { khi = 0; ++kdi; }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // hi*hi and hi*lo share their A operand: ONE MMA of width 2N against the adjacent [B_hi | B_lo] tiles writes main (columns
        // 0..N) and the first cross term (N..2N) at once; lo*hi then accumulates onto N..2N.
        // Two accumulators per tile: hi*hi in `tmem_d`, the two small cross terms in `tmem_x`.  The tensor core adds into the fp32
        // accumulator with truncation (round toward zero), a bias of ~0.5 ulp of the accumulator per MMA; keeping the 2/3 of the
        // MMAs that carry 2^-11-sized terms out of the main accumulator cuts that bias 3x (the cross accumulator is ~2^-10 of the
        // main one, its ulp is negligible).
        // All 32 lanes run the loops; the elected lane issues the MMAs and the commits (tc_elect_one(): the descriptors stay in
        // uniform registers, 8 UTCHMMA back to back per k-block instead of ~45 single-thread instructions per K-step).
        const bool leader = tc_elect_one();
        const uint32_t idesc = tc_idesc<F16>(kCvBM, N), idesc2 = tc_idesc<F16>(kCvBM, 2 * N);
        const uint64_t desc0 = tc_smem_desc(smem_u32(tiles));
        int st = 0, acc = 0, cbm = 0;
        uint32_t phs = 0, acc_ph = 0;
        if (p.khv) {
            const uint64_t bring = tc_desc_add(desc0, (uint32_t)p.a_slots * 2u * p.a_part);
            const uint32_t view = (uint32_t)p.bw * 128u;          // bytes per halo row: tap kh starts kh rows further down
            int sa_i = 0;
            uint32_t pha = 0;
            const int ngroups = p.kd * 3 * p.ncb;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_ph ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 2 * N);
                const uint32_t tmem_x = tmem_d + (uint32_t)N;
                for (int g = 0; g < ngroups; ++g) {
                    mbar_wait(&fullA[sa_i], pha);
                    tc_fence_after();
                    const uint64_t sa = tc_desc_add(desc0, (uint32_t)sa_i * 2u * p.a_part);
                    const int ksteps = (cbm == p.ncb - 1) ? p.klast : 4;    // zero-padded tail of the last channel block
                    if (++cbm == p.ncb) cbm = 0;
                    for (int khi = 0; khi < 3; ++khi) {
                        mbar_wait(&full_bar[st], phs);
                        tc_fence_after();
                        const uint64_t b_hi0 = tc_desc_add(bring, (uint32_t)st * 2u * b_part);
                        const uint64_t a_hi0 = tc_desc_add(sa, (uint32_t)khi * view), a_lo0 = tc_desc_add(a_hi0, p.a_part);
                        if (leader) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k >= ksteps) break;
                                tc_mma<F16>(tmem_d, tc_desc_add(a_hi0, k * 32), tc_desc_add(b_hi0, k * 32), idesc2, (g | khi | k) != 0 ? 1u : 0u);
                                tc_mma<F16>(tmem_x, tc_desc_add(a_lo0, k * 32), tc_desc_add(b_hi0, k * 32), idesc, 1u);
                            }
                            tc_commit(&empty_bar[st]);
                        }
                        if (++st == p.b_slots) { st = 0; phs ^= 1u; }
                    }
                    if (leader) tc_commit(&emptyA[sa_i]);
                    if (++sa_i == p.a_slots) { sa_i = 0; pha ^= 1u; }
                }
                if (leader) tc_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
            }
        } else {
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_ph ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * 2 * N);
                const uint32_t tmem_x = tmem_d + (uint32_t)N;
                for (int kb = 0; kb < p.nkb; ++kb) {
                    mbar_wait(&full_bar[st], phs);
                    tc_fence_after();
                    const uint64_t a_hi0 = tc_desc_add(desc0, (uint32_t)st * stage_bytes);
                    const uint64_t a_lo0 = tc_desc_add(a_hi0, kCvATile), b_hi0 = tc_desc_add(a_hi0, 2 * kCvATile);
                    const int ksteps = (cbm == p.ncb - 1) ? p.klast : 4;
                    if (++cbm == p.ncb) cbm = 0;
                    if (leader && !(p.dbg & 16)) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k >= ksteps) break;
                            tc_mma<F16>(tmem_d, tc_desc_add(a_hi0, k * 32), tc_desc_add(b_hi0, k * 32), idesc2, (kb | k) != 0 ? 1u : 0u);   // [hi*hi | hi*lo]
                            tc_mma<F16>(tmem_x, tc_desc_add(a_lo0, k * 32), tc_desc_add(b_hi0, k * 32), idesc, 1u);                      // + lo*hi
                        }
                    }
                    if (leader) tc_commit(&empty_bar[st]);
                    if (++st == stages) { st = 0; phs ^= 1u; }
                }
                if (leader) tc_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue =================
        // Low-K layers (DLA levels: 9 k-blocks per tile) are bound by this part, not by the MMAs (ncu: 34 % of the samples waited
        // on the residual loads, tensor pipe 25 % active).  8 warps share a tile -- warps w and w + 4 read the same TMEM lane
        // group (lanes 32 (w % 4) ..) and take alternating 16-column chunks -- and a warp requests the residual rows of its chunks
        // BEFORE it waits for the accumulator.
        const int lg = warp & 3;                 // TMEM lane group this warp may read
        const int chalf = (warp - 2) >> 2;       // which of the two warps of the lane group
        if (p.stg) {
            // ---- staged epilogue (2-D layers: few k-blocks per tile, the epilogue is what bounds them) ----
            // With a thread per accumulator row, every global access of a warp touches 32 different 128-byte lines (32 L1
            // wavefronts per instruction; ncu: the LSU, not the tensor pipe or DRAM, was the busy unit on the DLA layers).  Here a
            // warp owns 32 rows x 32 columns: it dumps the raw accumulator sums into its private staging rows, and then walks the
            // region with lanes ALONG the channels (8 lanes = 128 contiguous bytes of one output row): BatchNorm, residual, ReLU,
            // the operand split and all global loads / stores happen in that coalesced phase (4 wavefronts per instruction).
            float *stg = reinterpret_cast<float *>(tiles + p.stg_off) + (size_t)(warp - 2) * (32 * kCvStgRowF);
            const int sr = lane >> 3, sc4 = (lane & 7) * 4;
            int acc = 0;
            uint32_t acc_ph = 0;
            float amax = 0.f;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int mt = tile / p.n_ntiles, nt = tile - mt * p.n_ntiles;
                int n, d0, h0, w0;
                conv_tile_origin(p, mt, n, d0, h0, w0);
                size_t rowoff[8];                        // float offset of the first channel of the 8 rows this lane stores
                uint32_t rin = 0;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int mm = lg * 32 + it * 4 + sr;
                    const int ww = mm % p.bw, r1 = mm / p.bw, hh = r1 % p.bh, dd = r1 / p.bh;
                    if (d0 + dd < p.D) rin |= 1u << it;
                    rowoff[it] = ((((size_t)n * p.D + d0 + dd) * p.H + h0 + hh) * p.W + w0 + ww) * (size_t)p.ldy + (size_t)nt * N;
                }
                const int cbase = nt * N;
                const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * 2 * N);
                bool waited = false;
                for (int c0 = 32 * chalf; ; c0 += 64) {
                    const bool work = c0 < N && !(p.dbg & 8);
                    const bool active = work && c0 + sc4 < N;
                    float4 rres[8];
                    if (active && p.residual) {
#pragma unroll
                        for (int it = 0; it < 8; ++it)
                            if (rin >> it & 1) rres[it] = __ldg(reinterpret_cast<const float4 *>(p.residual + rowoff[it] + c0 + sc4));
                    }
                    if (!waited) {
                        mbar_wait(&tmem_full[acc], acc_ph);
                        tc_fence_after();
                        waited = true;
                    }
                    if (!work) {
                        if (32 * chalf >= N || (p.dbg & 8)) {   // a warp without columns (N <= 32) still hands the accumulator back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                        }
                        break;
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int c = c0 + 16 * h;
                        if (c < N) {
                            float v[16], vx[16];
                            tc_ld16(taddr + (uint32_t)c, v);
                            tc_ld16(taddr + (uint32_t)(N + c), vx);
                            float4 *sp = reinterpret_cast<float4 *>(stg + lane * kCvStgRowF + 16 * h);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float4 o;
                                o.x = F16 ? fmaf(vx[4 * j], kF16LoInv, v[4 * j] * p.acc_fix) : fmaf(v[4 * j], p.acc_fix, vx[4 * j]);
                                o.y = F16 ? fmaf(vx[4 * j + 1], kF16LoInv, v[4 * j + 1] * p.acc_fix) : fmaf(v[4 * j + 1], p.acc_fix, vx[4 * j + 1]);
                                o.z = F16 ? fmaf(vx[4 * j + 2], kF16LoInv, v[4 * j + 2] * p.acc_fix) : fmaf(v[4 * j + 2], p.acc_fix, vx[4 * j + 2]);
                                o.w = F16 ? fmaf(vx[4 * j + 3], kF16LoInv, v[4 * j + 3] * p.acc_fix) : fmaf(v[4 * j + 3], p.acc_fix, vx[4 * j + 3]);
                                sp[j] = o;
                            }
                        }
                    }
                    if (c0 + 64 >= N) {                  // last region of this warp: the accumulator is free again
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    }
                    __syncwarp();
                    if (active) {
                        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + cbase + c0 + sc4);
                        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + cbase + c0 + sc4);
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            if (!(rin >> it & 1)) continue;
                            float4 v = *reinterpret_cast<const float4 *>(stg + (it * 4 + sr) * kCvStgRowF + sc4);
                            v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                            if (p.relu == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            if (p.residual) { v.x += rres[it].x; v.y += rres[it].y; v.z += rres[it].z; v.w += rres[it].w; }
                            if (p.relu == 2) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            const size_t o = rowoff[it] + c0 + sc4;
                            if (p.y) *reinterpret_cast<float4 *>(p.y + o) = v;
                            if (F16 && p.y_hi) {
                                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                                uint32_t h0, l0, h1, l1;
                                f16_split2(v.x, v.y, h0, l0);
                                f16_split2(v.z, v.w, h1, l1);
                                *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(p.y_hi) + o) = make_uint2(h0, h1);
                                *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(p.y_lo) + o) = make_uint2(l0, l1);
                            } else if (p.y_hi) {
                                float4 h, l;
                                h.x = tf32_hi(v.x); l.x = v.x - h.x;
                                h.y = tf32_hi(v.y); l.y = v.y - h.y;
                                h.z = tf32_hi(v.z); l.z = v.z - h.z;
                                h.w = tf32_hi(v.w); l.w = v.w - h.w;
                                *reinterpret_cast<float4 *>(p.y_hi + o) = h;
                                *reinterpret_cast<float4 *>(p.y_lo + o) = l;
                            }
                        }
                    }
                    __syncwarp();                        // the staging rows are rewritten by the next region / tile
                }
                if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
            }
            if (F16 && p.rs && p.y_hi) range_commit(p.rs, amax);
        } else {
        const int m = lg * 32 + lane;
        int acc = 0;
        uint32_t acc_ph = 0;
        float amax = 0.f;                        // max |x| of what this thread splits into fp16 pairs (range guard)
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const int mt = tile / p.n_ntiles, nt = tile - mt * p.n_ntiles;
            int n, d0, h0, w0;
            conv_tile_origin(p, mt, n, d0, h0, w0);
            const int ww = m % p.bw, r1 = m / p.bw, hh = r1 % p.bh, dd = r1 / p.bh;
            const bool inb = d0 + dd < p.D;          // depth boxes may overhang the sample (2 images in boxes of 4 at the 12x40 level)
            const size_t vox = (((size_t)n * p.D + d0 + dd) * p.H + h0 + hh) * p.W + w0 + ww;
            const size_t row = vox * (size_t)p.ldy + (size_t)nt * N;     // float offset of this row's first channel
            const size_t orow = !p.pool ? row
                                        : ((((size_t)n * p.D + d0 + dd) * (p.H >> 1) + ((h0 + hh) >> 1)) * (p.W >> 1) + ((w0 + ww) >> 1)) *
                                                  (size_t)p.ldy + (size_t)nt * N;       // pooled output voxel
            float4 rres[4][4];                       // residual of this warp's chunks (N <= 128: 8 chunks over two warps)
            if (p.residual && inb) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = (2 * q + chalf) * 16;
                    if (c < N) {
                        const float4 *rp = reinterpret_cast<const float4 *>(p.residual + row + c);
#pragma unroll
                        for (int j = 0; j < 4; ++j) rres[q][j] = __ldg(rp + j);
                    }
                }
            }
            mbar_wait(&tmem_full[acc], acc_ph);
            tc_fence_after();
            const int cbase = nt * N;
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * 2 * N);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = (2 * q + chalf) * 16;
                if (c >= N || (p.dbg & 8)) break;
                float v[16], vx[16];
                tc_ld16(taddr + (uint32_t)c, v);
                tc_ld16(taddr + (uint32_t)(N + c), vx);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float sum = F16 ? fmaf(vx[j], kF16LoInv, v[j] * p.acc_fix) : fmaf(v[j], p.acc_fix, vx[j]);
                    float o = fmaf(sum, s_scale[cbase + c + j], s_shift[cbase + c + j]);
                    if (p.relu == 1) o = fmaxf(o, 0.f);
                    v[j] = o;
                }
                if (p.residual && inb) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 rr = rres[q][j];
                        v[4 * j] += rr.x; v[4 * j + 1] += rr.y; v[4 * j + 2] += rr.z; v[4 * j + 3] += rr.w;
                    }
                }
                if (p.relu == 2) {                                   // BasicBlock: relu(bn(conv) + residual)
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (p.pool) {
                    // 2x2 window over (h, w): the four voxels are lanes l, l^1 (next column) and l^bw, l^bw^1 (next row) of this
                    // warp (bw <= 16, even tile rows); the even-column / even-row lane stores the maximum
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
                        v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], p.bw));
                    }
                }
                const bool writer = inb && (!p.pool || !((ww | hh) & 1));
                if (writer && p.y) {
                    float4 *yp = reinterpret_cast<float4 *>(p.y + orow + c);
#pragma unroll
                    for (int j = 0; j < 4; ++j) yp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if (F16 && writer && p.y_hi) {
                    uint32_t hh[8], ll[8];
#pragma unroll
                    for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(v[j]));
#pragma unroll
                    for (int j = 0; j < 8; ++j) f16_split2(v[2 * j], v[2 * j + 1], hh[j], ll[j]);
                    uint4 *hp = reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(p.y_hi) + orow + c);
                    uint4 *lp = reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(p.y_lo) + orow + c);
                    hp[0] = make_uint4(hh[0], hh[1], hh[2], hh[3]); hp[1] = make_uint4(hh[4], hh[5], hh[6], hh[7]);
                    lp[0] = make_uint4(ll[0], ll[1], ll[2], ll[3]); lp[1] = make_uint4(ll[4], ll[5], ll[6], ll[7]);
                } else if (writer && p.y_hi) {
                    float4 *hp = reinterpret_cast<float4 *>(p.y_hi + orow + c);
                    float4 *lp = reinterpret_cast<float4 *>(p.y_lo + orow + c);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 h, l;
                        h.x = tf32_hi(v[4 * j]);     l.x = v[4 * j] - h.x;
                        h.y = tf32_hi(v[4 * j + 1]); l.y = v[4 * j + 1] - h.y;
                        h.z = tf32_hi(v[4 * j + 2]); l.z = v[4 * j + 2] - h.z;
                        h.w = tf32_hi(v[4 * j + 3]); l.w = v[4 * j + 3] - h.w;
                        hp[j] = h;
                        lp[j] = l;
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
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
        else
            cudaGetLastError();
    }
    return fn;
}

// channels-last activation [Nn, D, H, W, C] fp32 -> box {32, bw, bh, bd, 1}, 128-byte swizzle, zero fill out of bounds
static int make_act_tmap(CUtensorMap *tm, const float *base, int Nn, int D, int H, int W, int C, int bd, int bh, int bw,
                         int sh = 1, int sw = 1, int halo_h = 0, bool f16 = false)
{
    PFN_cuTensorMapEncodeTiled_v12000 enc = encode_fn();
    if (!enc) {
        set_error("conv_tc: cuTensorMapEncodeTiled is not available from the driver");
        return SIDE_ERR_CUDA;
    }
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)Nn};
    const cuuint64_t es_b = f16 ? 2 : 4;       // fp16 activations: 64 channels per 128-byte box row instead of 32
    cuuint64_t strides[4] = {(cuuint64_t)C * es_b, (cuuint64_t)W * C * es_b, (cuuint64_t)H * W * C * es_b,
                             (cuuint64_t)D * H * W * C * es_b};
    // strided convolution: the box spans bw*sw x bh*sh input pixels, traversed with element strides (sw, sh), i.e. it still
    // delivers bw x bh pixels -- the ones a stride-s convolution reads for bw x bh outputs
    cuuint32_t box[5] = {f16 ? 64u : 32u, (cuuint32_t)(bw * sw), (cuuint32_t)(bh * sh + 2 * halo_h), (cuuint32_t)bd, 1};
    cuuint32_t es[5] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1, 1};
    CUresult r = enc(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float *>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("conv_tc: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
        return SIDE_ERR_CUDA;
    }
    return SIDE_OK;
}

// exported to conv_tct.cu
int conv_make_act_tmap(CUtensorMap *tm, const float *base, int Nn, int D, int H, int W, int C, int bd, int bh, int bw, int sh,
                       int sw, int halo_h, bool f16)
{
    return make_act_tmap(tm, base, Nn, D, H, W, C, bd, bh, bw, sh, sw, halo_h, f16);
}
bool conv_tct_supported(int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int stride, bool f16);
int conv_tct_launch(const float *x_hi, const float *x_lo, const float *wp, const float *scale, const float *shift,
                    const float *residual, float *y, float *y_hi, float *y_lo, int Nn, int D, int H, int W, int Cin, int kd,
                    int relu, int sm_count, cudaStream_t st, bool f16);

// fp16 weight tiles: w [Nt rows of this n-tile][Cin][taps] -> [k-block of 64 channels][hi|lo'][Nt x 64 halves], swizzled
__global__ void __launch_bounds__(256) conv_tc_weight_prep_f16_kernel(const float *__restrict__ w, __half *__restrict__ wp, int Nt,
                                                                     int Cin, int taps)
{
    const long long total = (long long)taps * Cin * Nt;
    const int ncb = Cin / 64;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i % 64);
        const int n = (int)((i / 64) % Nt);
        const int kb = (int)(i / ((long long)64 * Nt));
        const int tap = kb / ncb, cb = kb - tap * ncb;
        const float v = __ldg(w + ((size_t)n * Cin + cb * 64 + kk) * taps + tap);
        const size_t tile = (size_t)kb * 2 * Nt * 64;
        const uint32_t off = (sw128(n, kk >> 3) >> 1) + (kk & 7);
        __half h, l;
        f16_split(v, h, l);
        wp[tile + off] = h;
        wp[tile + (size_t)Nt * 64 + off] = l;
    }
}

static int g_sm_count = 0;
static int g_disable_tct = 0;
static int g_enable_khv = 0;      // test hook: side_conv_tc_set_mode
static int g_dbg = 0;

}  // namespace side

using namespace side;

// n-tile width: all of Cout when it fits one accumulator pair (<= 128), otherwise 128-wide tiles
static int conv_ntile(int Cout) { return Cout <= 128 ? Cout : 128; }

extern "C" size_t side_conv_tc_weight_bytes(int Cin, int Cout, int taps)
{
    if (Cin <= 0 || Cout <= 0 || taps <= 0) return 0;
    return sizeof(float) * 2 * (size_t)Cin * Cout * taps;
}

static int conv_tc_prep_weights_impl(const float *w, float *wp, int Cout, int Cin, int taps, void *stream, bool f16);
extern "C" int side_conv_tc_prep_weights(const float *w, float *wp, int Cout, int Cin, int taps, void *stream)
{
    return conv_tc_prep_weights_impl(w, wp, Cout, Cin, taps, stream, false);
}
/* fp16 "3xFP16" tiles (half the bytes of the tf32 ones: side_conv_tc_weight_bytes(...) / 2); needs Cin % 64 == 0 */
extern "C" int side_conv_tc_prep_weights_f16(const float *w, void *wp, int Cout, int Cin, int taps, void *stream)
{
    return conv_tc_prep_weights_impl(w, reinterpret_cast<float *>(wp), Cout, Cin, taps, stream, true);
}
static int conv_tc_prep_weights_impl(const float *w, float *wp, int Cout, int Cin, int taps, void *stream, bool f16)
{
    SIDE_REQUIRE(!f16 || Cin % 64 == 0, "side_conv_tc_prep_weights_f16: needs Cin %% 64 == 0");
    SIDE_REQUIRE(Cin > 0 && Cin % 32 == 0 && Cout >= 16 && Cout % 16 == 0 && Cout <= kCvMaxCout && taps > 0 &&
                     (Cout <= 128 || Cout % 128 == 0),
                 "side_conv_tc_prep_weights: needs Cin %% 32 == 0 and Cout %% 16 == 0 (<= 128) or Cout %% 128 == 0 (<= %d)",
                 kCvMaxCout);
    SIDE_REQUIRE_DEV(w); SIDE_REQUIRE_DEV(wp);
    // one swizzled tile set per n-tile: [n_tile][k-block][hi|lo][Nt x 32]
    const int Nt = conv_ntile(Cout);
    if (f16) {
        __half *wh = reinterpret_cast<__half *>(wp);
        for (int nt = 0; nt < Cout / Nt; ++nt) {
            const long long total = (long long)Nt * Cin * taps;
            conv_tc_weight_prep_f16_kernel<<<(unsigned)std::min<long long>(1184, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
                w + (size_t)nt * Nt * Cin * taps, wh + (size_t)nt * 2 * Nt * Cin * taps, Nt, Cin, taps);
            SIDE_LAUNCH_CHECK("conv_tc_weight_prep_f16_kernel");
        }
        return SIDE_OK;
    }
    for (int nt = 0; nt < Cout / Nt; ++nt) {
        int rc = launch_tc_weight_prep(w + (size_t)nt * Nt * Cin * taps, wp + (size_t)nt * 2 * Nt * Cin * taps, Nt, Cin, taps, 1,
                                       (cudaStream_t)stream);
        if (rc) return rc;
    }
    return SIDE_OK;
}

static int conv3d_tc_fwd_impl(const float *x_hi, const float *x_lo, const float *wp, const float *scale, const float *shift,
                              const float *residual, float *y, float *y_hi, float *y_lo, int Nn, int D, int H, int W, int Cin,
                              int Cout, int kd, int kh, int kw, int stride_hw, int relu, void *stream, bool f16);
extern "C" int side_conv3d_tc_fwd(const float *x_hi, const float *x_lo, const float *wp, const float *scale,
                                  const float *shift, const float *residual, float *y, float *y_hi, float *y_lo, int Nn,
                                  int D, int H, int W, int Cin, int Cout, int kd, int kh, int kw, int stride_hw, int relu,
                                  void *stream)
{
    return conv3d_tc_fwd_impl(x_hi, x_lo, wp, scale, shift, residual, y, y_hi, y_lo, Nn, D, H, W, Cin, Cout, kd, kh, kw, stride_hw,
                              relu, stream, false);
}
/* Same convolution on kind::f16 MMAs ("3xFP16": x = hi + lo' * 2^-11, fp16 pairs; twice the tensor rate of 3xTF32, half the
 * operand bytes).  x_hi / x_lo / y_hi / y_lo are fp16 arrays, wp comes from side_conv_tc_prep_weights_f16; y, residual, scale
 * and shift stay fp32.  Needs Cin % 64 == 0. */
extern "C" int side_conv3d_tc_fwd_f16(const void *x_hi, const void *x_lo, const void *wp, const float *scale, const float *shift,
                                      const float *residual, float *y, void *y_hi, void *y_lo, int Nn, int D, int H, int W,
                                      int Cin, int Cout, int kd, int kh, int kw, int stride_hw, int relu, void *stream)
{
    return conv3d_tc_fwd_impl(reinterpret_cast<const float *>(x_hi), reinterpret_cast<const float *>(x_lo),
                              reinterpret_cast<const float *>(wp), scale, shift, residual, y, reinterpret_cast<float *>(y_hi),
                              reinterpret_cast<float *>(y_lo), Nn, D, H, W, Cin, Cout, kd, kh, kw, stride_hw, relu, stream, true);
}
static int conv3d_tc_fwd_impl(const float *x_hi, const float *x_lo, const float *wp, const float *scale, const float *shift,
                              const float *residual, float *y, float *y_hi, float *y_lo, int Nn, int D, int H, int W, int Cin,
                              int Cout, int kd, int kh, int kw, int stride_hw, int relu, void *stream, bool f16)
{
    // fp16: the last 64-channel block may be half full (Cin % 64 == 32): the TMA box reads past the channel extent (zero fill),
    // the weight tiles are zero-padded by side_conv_tc_prep_weights_f16's caller, and the MMAs of the empty half are skipped
    SIDE_REQUIRE(Nn >= 0 && D > 0 && H > 0 && W > 0, "side_conv3d_tc_fwd: bad shape");
    SIDE_REQUIRE(Cin > 0 && Cin % 32 == 0 && Cout >= 16 && Cout % 16 == 0 && Cout <= kCvMaxCout && (Cout <= 128 || Cout % 128 == 0),
                 "side_conv3d_tc_fwd: needs Cin %% 32 == 0 and Cout %% 16 == 0 (<= 128) or Cout %% 128 == 0 (got %d -> %d)", Cin, Cout);
    SIDE_REQUIRE((kd == 3 && kh == 3 && kw == 3) || (kd == 1 && kh == 3 && kw == 3) || (kd == 1 && kh == 1 && kw == 1),
                 "side_conv3d_tc_fwd: kernel must be 3x3x3, 1x3x3 or 1x1x1");
    SIDE_REQUIRE(stride_hw == 1 || (stride_hw == 2 && kd == 1 && H % 2 == 0 && W % 2 == 0),
                 "side_conv3d_tc_fwd: stride must be 1, or 2 for 2-D kernels on even maps");
    const int pool = (relu >> 2) & 1;              // bit 2: fused MaxPool3d((1,2,2))
    relu &= 3;
    SIDE_REQUIRE(relu >= 0 && relu <= 2, "side_conv3d_tc_fwd: relu must be 0 (none), 1 (before the residual) or 2 (after it)");
    if (Nn == 0) return SIDE_OK;
    const int Ho = H / stride_hw, Wo = W / stride_hw;          // "same" padding: ceil(H / s) with even H
    // 128-voxel box of OUTPUT voxels: bw = largest power of two <= 128 dividing Wo, then rows, then depth slices
    int bw = 128;
    while (bw > 1 && Wo % bw) bw >>= 1;
    int bh = 128 / bw;
    while (bh > 1 && Ho % bh) bh >>= 1;
    const int bd = 128 / (bw * bh);
    SIDE_REQUIRE(bd >= 1 && (D % bd == 0 || kd == 1), "side_conv3d_tc_fwd: %dx%dx%d does not tile into 128-voxel boxes (box %dx%dx%d)", D, Ho, Wo,
                 bd, bh, bw);       // 2-D kernels (D = batch): a last box hanging over the batch is zero-filled / skipped
    SIDE_REQUIRE(bw * stride_hw <= 256 && bh * stride_hw <= 256, "side_conv3d_tc_fwd: box too large for the stride");
    SIDE_REQUIRE((long long)Nn * D * H * W < (1ll << 31), "side_conv3d_tc_fwd: too many voxels");
    SIDE_REQUIRE(y || (y_hi && y_lo), "side_conv3d_tc_fwd: no output requested");
    SIDE_REQUIRE((y_hi == nullptr) == (y_lo == nullptr), "side_conv3d_tc_fwd: y_hi and y_lo go together");
    SIDE_REQUIRE_DEV(x_hi); SIDE_REQUIRE_DEV(x_lo); SIDE_REQUIRE_DEV(wp);
    if (y) SIDE_REQUIRE_DEV(y);
    if (y_hi) { SIDE_REQUIRE_DEV(y_hi); SIDE_REQUIRE_DEV(y_lo); }

    // kh-view mode when the geometry allows it and the two rings fit in shared memory
    const uint32_t b_slot = 2 * (uint32_t)conv_ntile(Cout) * 128u;
    const uint32_t a_part = (uint32_t)(bh + 2) * bw * 128u;
    int khv = (kh == 3 && stride_hw == 1 && bd == 1 && (bw % 8) == 0 && g_enable_khv) ? 1 : 0;
    int b_slots = 0;
    if (khv) {
        const uint32_t budget = 196u * 1024u - kCvStgBytes;          // room for the staging area of the coalescing epilogue
        b_slots = 2u * 2u * a_part < budget ? std::min(4, (int)((budget - 2u * 2u * a_part) / b_slot)) : 0;
        if (b_slots < 2) khv = 0;
    }
    int rc;
    if (g_sm_count == 0) {
        int dev = 0;
        SIDE_CUDA(cudaGetDevice(&dev));
        SIDE_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    // Cout == 64 on 256-voxel slices: role-swapped kernel (weights on M, 256 voxels on N), see conv_tct.cu
    SIDE_REQUIRE(!pool || (stride_hw == 1 && bw >= 2 && bw <= 16 && bh % 2 == 0 && H % 2 == 0 && W % 2 == 0),
                 "side_conv3d_tc_fwd: the fused 2x2 max-pool needs stride 1, even H and W and boxes of 2..16 columns x an even number of rows");
    if (!pool && !g_disable_tct && !g_dbg && conv_tct_supported(D, H, W, Cin, Cout, kd, kh, kw, stride_hw, f16))   // (takes Cin % 32 == 0)
        return conv_tct_launch(x_hi, x_lo, wp, scale, shift, residual, y, y_hi, y_lo, Nn, D, H, W, Cin, kd, relu, g_sm_count,
                               (cudaStream_t)stream, f16);
    CUtensorMap tm_hi, tm_lo;
    if ((rc = make_act_tmap(&tm_hi, x_hi, Nn, D, H, W, Cin, bd, bh, bw, stride_hw, stride_hw, khv, f16))) return rc;
    if ((rc = make_act_tmap(&tm_lo, x_lo, Nn, D, H, W, Cin, bd, bh, bw, stride_hw, stride_hw, khv, f16))) return rc;

    ConvTcParams p;
    p.wp = wp; p.y = y; p.y_hi = y_hi; p.y_lo = y_lo; p.scale = scale; p.shift = shift; p.residual = residual;
    p.relu = relu; p.N = conv_ntile(Cout); p.Ntot = Cout; p.n_ntiles = Cout / p.N; p.ldy = Cout;
    p.ncb = f16 ? (Cin + 63) / 64 : Cin / 32; p.nkb = kd * kh * kw * p.ncb;
    p.klast = (f16 && Cin % 64 == 32) ? 2 : 4;
    p.pool = pool;
    const int dt = (D + bd - 1) / bd;
    const long long mtiles = (long long)Nn * dt * (Ho / bh) * (Wo / bw);
    SIDE_REQUIRE(mtiles * p.n_ntiles < (1ll << 31), "side_conv3d_tc_fwd: too many tiles");
    p.ntiles = (int)(mtiles * p.n_ntiles);
    p.D = D; p.H = Ho; p.W = Wo; p.bw = bw; p.bh = bh; p.bd = bd; p.wt = Wo / bw; p.ht = Ho / bh; p.dt = dt;
    p.kd = kd; p.kh = kh; p.kw = kw; p.sh = stride_hw; p.sw = stride_hw;
    const uint32_t stage_bytes = 2 * kCvATile + 2 * (uint32_t)p.N * 128u;
    // 2-D layers (few k-blocks per tile): coalescing epilogue staged through 36 KB of shared memory behind the operand ring
    static const int stg_env = [] { const char *e = getenv("SIDE_CONV_STG"); return e ? atoi(e) : 1; }();
    // (measured per layer shape, B = 16: 64->64 @ 96x320 194 -> 167 us, 64->128 @ 48x160 100 -> 71 us, 128->128 @ 48x160 128 -> 113 us;
    // from 36 k-blocks per tile on, the MMAs hide the epilogue and the ring depth given up for the staging area costs 3 %)
    p.stg = (stg_env && !pool && kd == 1 && p.nkb <= 18 && p.N >= 64) ? 1 : 0;
    const uint32_t ring_budget = 196u * 1024u - (p.stg ? kCvStgBytes : 0u);
    p.stages = std::max(2, std::min((int)(ring_budget / stage_bytes), kCvMaxStages));
    p.stg_off = khv ? 2u * 2u * a_part + (uint32_t)b_slots * b_slot : (uint32_t)p.stages * stage_bytes;
    p.khv = khv; p.a_slots = 2; p.b_slots = b_slots; p.a_part = a_part; p.dbg = g_dbg;
    p.rs = f16 ? range_slot_next() : nullptr;
    p.acc_fix = tc_acc_fix(kd * kh * kw * ((p.ncb - 1) * 4 + p.klast));
    const size_t smem = (khv ? (size_t)2 * 2 * a_part + (size_t)b_slots * b_slot : (size_t)p.stages * stage_bytes) + (p.stg ? kCvStgBytes : 0u) + 1024;
    if ((rc = set_smem_attr(f16 ? (const void *)conv_tc_kernel<true> : (const void *)conv_tc_kernel<false>, smem))) return rc;
    if (g_sm_count == 0) {
        int dev = 0;
        SIDE_CUDA(cudaGetDevice(&dev));
        SIDE_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const unsigned grid = (unsigned)std::min(p.ntiles, g_sm_count);
    if (f16) conv_tc_kernel<true><<<grid, kCvThreads, smem, (cudaStream_t)stream>>>(tm_hi, tm_lo, p);
    else conv_tc_kernel<false><<<grid, kCvThreads, smem, (cudaStream_t)stream>>>(tm_hi, tm_lo, p);
    SIDE_LAUNCH_CHECK("conv_tc_kernel");
    return SIDE_OK;
}

// Plain GEMM on the same kernel (1x1x1 "convolution" over rows/128 x 128 "voxels"): y[r][n] = sum_k x[r][k] w[n][k] for a
// block of Ncols columns of a matrix with row stride ldy.  x is pre-split (hi, lo), wp holds Ncols/Nt swizzled n-tiles
// ([n-tile][k-block][hi|lo][Nt x 32]).  Used by the DCN backward (dcn_bwd_cl.cu).
namespace side {
int conv_tc_rows_gemm(const float *x_hi, const float *x_lo, const float *wp, float *y, long long ldy, long long rows, int K,
                      int Ncols, int Nt, cudaStream_t st)
{
    SIDE_REQUIRE(rows > 0 && rows % kCvBM == 0 && rows / kCvBM < (1ll << 31) && K > 0 && K % 32 == 0,
                 "conv_tc_rows_gemm: rows %% 128 == 0 and K %% 32 == 0 required");
    SIDE_REQUIRE(Nt >= 16 && Nt % 16 == 0 && Nt <= 128 && Ncols % Nt == 0 && Ncols <= kCvMaxCout, "conv_tc_rows_gemm: bad column tiling");
    int rc;
    const int Hr = (int)(rows / kCvBM);
    CUtensorMap tm_hi, tm_lo;
    if ((rc = make_act_tmap(&tm_hi, x_hi, 1, 1, Hr, kCvBM, K, 1, 1, kCvBM))) return rc;
    if ((rc = make_act_tmap(&tm_lo, x_lo, 1, 1, Hr, kCvBM, K, 1, 1, kCvBM))) return rc;
    ConvTcParams p;
    p.wp = wp; p.y = y; p.y_hi = nullptr; p.y_lo = nullptr; p.scale = nullptr; p.shift = nullptr; p.residual = nullptr;
    p.relu = 0; p.N = Nt; p.Ntot = Ncols; p.n_ntiles = Ncols / Nt; p.ldy = ldy;
    p.ncb = K / 32; p.nkb = p.ncb; p.klast = 4; p.pool = 0;
    SIDE_REQUIRE((long long)Hr * p.n_ntiles < (1ll << 31), "conv_tc_rows_gemm: too many tiles");
    p.ntiles = Hr * p.n_ntiles;
    p.D = 1; p.H = Hr; p.W = kCvBM; p.bw = kCvBM; p.bh = 1; p.bd = 1; p.wt = 1; p.ht = Hr; p.dt = 1;
    p.kd = 1; p.kh = 1; p.kw = 1; p.sh = 1; p.sw = 1;
    const uint32_t stage_bytes = 2 * kCvATile + 2 * (uint32_t)p.N * 128u;
    p.stages = std::max(2, std::min((int)((196u * 1024u) / stage_bytes), kCvMaxStages));
    p.khv = 0; p.a_slots = 2; p.b_slots = 0; p.a_part = 0; p.dbg = 0; p.rs = nullptr; p.stg = 0; p.stg_off = 0;
    p.acc_fix = tc_acc_fix(p.ncb * 4);
    const size_t smem = (size_t)p.stages * stage_bytes + 1024;
    if ((rc = set_smem_attr((const void *)conv_tc_kernel<false>, smem))) return rc;
    if (g_sm_count == 0) {
        int dev = 0;
        SIDE_CUDA(cudaGetDevice(&dev));
        SIDE_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    conv_tc_kernel<false><<<(unsigned)std::min(p.ntiles, g_sm_count), kCvThreads, smem, st>>>(tm_hi, tm_lo, p);
    SIDE_LAUNCH_CHECK("conv_tc_kernel");
    return SIDE_OK;
}
}  // namespace side

/* test / benchmark hook (bit mask).  0 = default.  1: the voxel-major kernel reuses one halo box for the three vertical taps
 * (measured: no gain there -- that kernel is bound by the ~115-cycle floor of each tcgen05.mma, not by operand delivery);
 * 32: disable the role-swapped Cout = 64 kernel (conv_tct.cu); 2 / 4 / 8 / 16: timing experiments that skip the B copies /
 * A copies / epilogue / MMAs (results are garbage). */
extern "C" int side_conv_tc_set_mode(int mode)
{
    g_enable_khv = (mode & 1) && !(mode & 30);
    g_dbg = mode & 30;
    g_disable_tct = (mode & 32) != 0;
    return SIDE_OK;
}
