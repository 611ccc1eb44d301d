// dcn_gw_tc.cu -- weight gradient of the DCNv2 backward on tcgen05.
//
// Reference: the second Sgemm of dcn_v2_cuda_backward (DCNv2/src/cuda/dcn_v2_cuda.cu:306-316): grad_weight += grad_output_n *
// columns^T, one sample at a time.  Here:  gW[o][k'] = sum over (sample, pixel) of gy[b][o][p] * col[b*P + p][k'],  k' = tap*Cin + c.
//
// The reduction runs over PIXELS.  gy is NCHW ([o][p], pixels contiguous), so it is a K-major A operand as it lies in memory.
// The column buffer is pixel-major ([p][k'], k' contiguous, as the scatter kernel writes it): with the pixels as K that is an
// MN-MAJOR B operand -- tcgen05 reads it in place (instruction-descriptor bit 16, canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO))
// in 16-byte units under the 128-byte swizzle: a TMA box of 64 k' x 64 pixels is 8 swizzle atoms stacked along K, SBO = 1024 B,
// and the next 64 k' sit LBO = 8 KB further), so no transposed copy of the 283 MB buffer is needed.
// Operands are 3xFP16 pairs (hi, lo * 2^11, see tc_common.cuh): the scatter kernel emits the column pairs, a small kernel
// splits gy.  A CTA owns (128 output-channel rows, <= 256 columns, a range of 64-pixel k-blocks): one TMA warp, one MMA thread
// (hi*hi into the main accumulator, hi*lo' + lo'*hi into the cross accumulator: 512 TMEM columns), four epilogue warps that add
// the tile into the fp32 gradient with 16-byte reductions (split-K).  Cout = 64 uses the upper half of the tile as TMA zero fill.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>

#include "tc_common.cuh"

namespace side {

constexpr int kGwThreads = 192;
constexpr int kGwStages = 2;
constexpr uint32_t kGwAPart = 128 * 128;          // 16 KB: 128 rows (o) x 64 pixels fp16
constexpr uint32_t kGwBChunk = 64 * 128;          // 8 KB: 64 pixels (K rows) x 64 k' fp16
constexpr uint32_t kGwBPart = 4 * kGwBChunk;      // up to 256 columns
constexpr uint32_t kGwStage = 2 * kGwAPart + 2 * kGwBPart;   // 96 KB

struct GwParams {
    float *gw;                 // [Cout][Kp] fp32, accumulated into
    const uint32_t *gy_absmax; // bits of max |gy| over the whole batch (gw_absmax_kernel): fixes the power-of-two range scale
    int Cout, Kp, P, nb;       // P % 64 == 0
    int n_ntiles, n_mtiles, ksplits, kb_total;     // kb_total = nb * P / 64
    // fused convolution mode (conv_wgrad.cu): the B operand is the convolution INPUT read in place through shifted TMA boxes
    // -- no patch matrix.  A k-block = one box of bw x bh x bd = 64 output voxels; gy map [W, H, D, Cout, N], x map [C, W, H, D, N]
    int fused, bw, bh, bd, nbw, nbh, Cq, stride, kd, kh, kw;
};

__device__ __forceinline__ void gw_tma_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void gw_tma_5d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, int c4, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void gw_tma_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

// MN-major, SWIZZLE_128B operand: 64 MN elements (128 bytes) per swizzle row, 8-row atoms along K every SBO = 1024 bytes,
// the next 64 MN elements LBO bytes further
__device__ __forceinline__ uint64_t gw_desc_mn(uint32_t saddr, uint32_t lbo)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

// gy is a back-propagated gradient: its magnitude is arbitrary (1e-10 for mean-reduced losses over 96x320 maps, 1e+6 for summed
// ones), while fp16 pairs only resolve 2^-24 .. 65504.  The split therefore runs on gy * 2^s with s chosen from max |gy| so that
// the largest element lands in [2^12, 2^13): exact (power of two), 26 binades of headroom below before `hi` goes subnormal, and
// the weight-gradient epilogue multiplies the tile by 2^-s.  Both sides derive s from the same absmax word.
__device__ __forceinline__ float gw_range_scale(uint32_t absmax_bits, bool inverse)
{
    int e = (int)(absmax_bits >> 23) - 127;                 // floor(log2(max |gy|)); zero / subnormal maxima -> no scaling
    if (absmax_bits < 0x00800000u || absmax_bits >= 0x7F800000u) e = 12;
    int s = 12 - e;                                         // scale = 2^s
    s = max(-126, min(126, s));
    return __uint_as_float((uint32_t)(127 + (inverse ? -s : s)) << 23);
}

__global__ void __launch_bounds__(256) gw_absmax_kernel(const float4 *__restrict__ x, uint32_t *__restrict__ out, long long n4)
{
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));     // fmaxf drops NaNs
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));    // non-negative floats order like their bits
}

__global__ void __launch_bounds__(kGwThreads, 1) dcn_gw_tc_kernel(const __grid_constant__ CUtensorMap tm_gy_hi,
                                                                  const __grid_constant__ CUtensorMap tm_gy_lo,
                                                                  const __grid_constant__ CUtensorMap tm_col_hi,
                                                                  const __grid_constant__ CUtensorMap tm_col_lo, GwParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[kGwStages], empty_bar[kGwStages], tmem_full;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler as well
    unsigned char *tiles = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);

    // work item of this CTA
    int r = blockIdx.x;
    const int ks = r % p.ksplits; r /= p.ksplits;
    const int nt = r % p.n_ntiles;
    const int mt = r / p.n_ntiles;
    const int n0 = nt * 256, Nw = min(256, p.Kp - n0);              // multiple of 64
    const int per = (p.kb_total + p.ksplits - 1) / p.ksplits;
    const int kb0 = ks * per, kb1 = min(p.kb_total, kb0 + per);
    const int nkb = kb1 - kb0;
    const int kpi = (p.P + 63) / 64;                                 // k-blocks per sample; a partial last block is zero-filled
                                                                     // on the gy side (TMA), which cancels whatever rows B holds

    if (tid == 0) {
        for (int i = 0; i < kGwStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // all 32 lanes run the loop (uniform operands), the elected lane issues the copies -- see tc_elect_one()
        const bool leader = tc_elect_one();
        if (nkb > 0) {
            const uint32_t bytes = 2 * kGwAPart + 2 * (uint32_t)(Nw / 64) * kGwBChunk;
            int st = 0;
            uint32_t ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                const int bl = kb / kpi, p0 = (kb - bl * kpi) * 64;
                mbar_wait(&empty_bar[st], ph ^ 1u);
                unsigned char *sa = tiles + (size_t)st * kGwStage;
                if (!leader) {
                    if (++st == kGwStages) { st = 0; ph ^= 1u; }
                    continue;
                }
                mbar_expect_tx(&full_bar[st], bytes);
                if (p.fused) {
                    // box r of sample bl: output voxels (d0.., h0.., w0..); tap (td, th, tw) of column chunk j shifts the INPUT box
                    const int r = kb - bl * kpi, bx = r % p.nbw, by = (r / p.nbw) % p.nbh, bz = r / (p.nbw * p.nbh);
                    const int w0 = bx * p.bw, h0 = by * p.bh, d0 = bz * p.bd;
                    // gy pairs were written box-major by gw_split_gy_box_kernel: the 64 voxels of box r are contiguous (a multi-
                    // dimensional TMA box with an inner extent below 128 bytes is not laid out as one dense 128-byte K row)
                    gw_tma_3d(sa, &tm_gy_hi, r * 64, mt * 128, bl, &full_bar[st]);
                    gw_tma_3d(sa + kGwAPart, &tm_gy_lo, r * 64, mt * 128, bl, &full_bar[st]);
                    unsigned char *sbf = sa + 2 * kGwAPart;
                    for (int j = 0; j < Nw / 64; ++j) {
                        const int kq = n0 + 64 * j, tap = kq / p.Cq, c0 = kq - tap * p.Cq;
                        const int tw = tap % p.kw, th = (tap / p.kw) % p.kh, td = tap / (p.kw * p.kh);
                        const int xi = w0 * p.stride + tw - (p.kw - 1) / 2, yi = h0 * p.stride + th - (p.kh - 1) / 2,
                                  zi = d0 + td - (p.kd - 1) / 2;
                        gw_tma_5d(sbf + (size_t)j * kGwBChunk, &tm_col_hi, c0, xi, yi, zi, bl, &full_bar[st]);
                        gw_tma_5d(sbf + kGwBPart + (size_t)j * kGwBChunk, &tm_col_lo, c0, xi, yi, zi, bl, &full_bar[st]);
                    }
                    if (++st == kGwStages) { st = 0; ph ^= 1u; }
                    continue;
                }
                gw_tma_3d(sa, &tm_gy_hi, p0, mt * 128, bl, &full_bar[st]);                 // rows past Cout: zero fill
                gw_tma_3d(sa + kGwAPart, &tm_gy_lo, p0, mt * 128, bl, &full_bar[st]);
                unsigned char *sb = sa + 2 * kGwAPart;
                for (int j = 0; j < Nw / 64; ++j) {
                    gw_tma_2d(sb + (size_t)j * kGwBChunk, &tm_col_hi, n0 + 64 * j, bl * p.P + p0, &full_bar[st]);
                    gw_tma_2d(sb + kGwBPart + (size_t)j * kGwBChunk, &tm_col_lo, n0 + 64 * j, bl * p.P + p0, &full_bar[st]);
                }
                if (++st == kGwStages) { st = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // all 32 lanes run the loop, the elected lane issues the MMAs and the commits (tc_elect_one())
        const bool leader = tc_elect_one();
        if (nkb > 0) {
            // A: K-major (bit 15 = 0), B: MN-major (bit 16 = 1)
            const uint32_t idesc = tc_idesc_f16(128, Nw) | (1u << 16);
            const uint32_t tmem_d = tmem_base, tmem_x = tmem_base + 256u;
            const uint64_t adesc0 = tc_smem_desc(smem_u32(tiles)), bdesc0 = gw_desc_mn(smem_u32(tiles) + 2 * kGwAPart, kGwBChunk);
            int st = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(&full_bar[st], ph);
                tc_fence_after();
                const uint64_t sa = tc_desc_add(adesc0, (uint32_t)st * kGwStage), sb = tc_desc_add(bdesc0, (uint32_t)st * kGwStage);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                      // 16 pixels per instruction
                        const uint64_t a_hi = tc_desc_add(sa, k * 32), a_lo = tc_desc_add(sa, kGwAPart + k * 32);
                        const uint64_t b_hi = tc_desc_add(sb, k * 2048), b_lo = tc_desc_add(sb, kGwBPart + k * 2048);
                        tc_mma_f16(tmem_d, a_hi, b_hi, idesc, (i | k) != 0 ? 1u : 0u);
                        tc_mma_f16(tmem_x, a_hi, b_lo, idesc, (i | k) != 0 ? 1u : 0u);
                        tc_mma_f16(tmem_x, a_lo, b_hi, idesc, 1u);
                    }
                    tc_commit(&empty_bar[st]);
                }
                if (++st == kGwStages) { st = 0; ph ^= 1u; }
            }
            if (leader) tc_commit(&tmem_full);
        }
        __syncwarp();
    } else if (nkb > 0) {
        // epilogue: TMEM lane = output channel row, column = k'
        const int lg = warp & 3, o = mt * 128 + lg * 32 + lane;
        mbar_wait(&tmem_full, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16);
        float *gp = p.gw + (size_t)o * p.Kp + n0;
        const float inv = gw_range_scale(__ldg(p.gy_absmax), true), invx = inv * kF16LoInv;
        for (int c = 0; c < Nw; c += 16) {
            float v[16], vx[16];
            tc_ld16(taddr + (uint32_t)c, v);
            tc_ld16(taddr + 256u + (uint32_t)c, vx);
            if (o < p.Cout) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gp + c + 4 * j),
                                 "f"(fmaf(vx[4 * j], invx, v[4 * j] * inv)), "f"(fmaf(vx[4 * j + 1], invx, v[4 * j + 1] * inv)),
                                 "f"(fmaf(vx[4 * j + 2], invx, v[4 * j + 2] * inv)), "f"(fmaf(vx[4 * j + 3], invx, v[4 * j + 3] * inv))
                                 : "memory");
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// gy [B][Cout][P] fp32 -> fp16 pairs of (gy * 2^s), same layout
__global__ void __launch_bounds__(256) gw_split_gy_kernel(const float4 *__restrict__ x, uint2 *__restrict__ hi, uint2 *__restrict__ lo,
                                                         long long n4, const uint32_t *__restrict__ absmax)
{
    const float sc = gw_range_scale(__ldg(absmax), false);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(x + i);
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        uint32_t h0, l0, h1, l1;
        f16_split2(v.x, v.y, h0, l0);
        f16_split2(v.z, v.w, h1, l1);
        hi[i] = make_uint2(h0, h1);
        lo[i] = make_uint2(l0, l1);
    }
}

// gy [N][Cout][Do][Ho][Wo] fp32 -> fp16 pairs of (gy * 2^s) in BOX-MAJOR voxel order: position box * 64 + (dz * bh + dy) * bw + dx of
// every (n, o) row holds voxel (bz * bd + dz, by * bh + dy, bx * bw + dx), box = (bz * nbh + by) * nbw + bx
__global__ void __launch_bounds__(256) gw_split_gy_box_kernel(const float *__restrict__ x, __half *__restrict__ hi, __half *__restrict__ lo,
                                                             long long rows, int Do, int Ho, int Wo, int bw, int bh, int bd,
                                                             const uint32_t *__restrict__ absmax)
{
    const float sc = gw_range_scale(__ldg(absmax), false);
    const long long P = (long long)Do * Ho * Wo, total = rows * P / 2;
    const int nbw = Wo / bw, nbh = Ho / bh;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = (2 * i) / P;
        const int q = (int)(2 * i - row * P);                  // even position inside the row (bw is even: the pair shares a box row)
        const int box = q >> 6, r = q & 63, dx = r % bw, dy = (r / bw) % bh, dz = r / (bw * bh);
        const int bx = box % nbw, by = (box / nbw) % nbh, bz = box / (nbw * nbh);
        const size_t src = (size_t)row * P + ((size_t)(bz * bd + dz) * Ho + by * bh + dy) * Wo + bx * bw + dx;
        const float2 v = *reinterpret_cast<const float2 *>(x + src);
        uint32_t h, l;
        f16_split2(v.x * sc, v.y * sc, h, l);
        reinterpret_cast<uint32_t *>(hi)[i] = h;
        reinterpret_cast<uint32_t *>(lo)[i] = l;
    }
}

// 64-voxel box of the fused convolution weight gradient: bw x bh x bd with bw | Wo, bh | Ho, bd | Do (powers of two, bw even)
bool dcn_gw_box(int Do, int Ho, int Wo, int &bw, int &bh, int &bd)
{
    bw = 64;
    while (bw > 1 && Wo % bw) bw >>= 1;
    bh = 64 / bw;
    while (bh > 1 && Ho % bh) bh >>= 1;
    bd = 64 / (bw * bh);
    return bw >= 2 && Do % bd == 0 && bw * bh * bd == 64;
}

int dcn_gw_tc_split_gy_box(const float *gy, void *gy_pairs, int N, int Cout, int Do, int Ho, int Wo, int bw, int bh, int bd,
                           cudaStream_t st)
{
    const long long P = (long long)Do * Ho * Wo, n = (long long)N * Cout * P;
    __half *hi = reinterpret_cast<__half *>(gy_pairs);
    uint32_t *absmax = reinterpret_cast<uint32_t *>(hi + 2 * n);
    SIDE_CUDA(cudaMemsetAsync(absmax, 0, sizeof(uint32_t), st));
    const unsigned grid = (unsigned)std::min<long long>((n / 4 + 255) / 256, 148 * 16);
    gw_absmax_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(gy), absmax, n / 4);
    SIDE_LAUNCH_CHECK("gw_absmax_kernel");
    gw_split_gy_box_kernel<<<(unsigned)std::min<long long>((n / 2 + 255) / 256, 148 * 16), 256, 0, st>>>(gy, hi, hi + n, (long long)N * Cout,
                                                                                                     Do, Ho, Wo, bw, bh, bd, absmax);
    SIDE_LAUNCH_CHECK("gw_split_gy_box_kernel");
    return SIDE_OK;
}

static PFN_cuTensorMapEncodeTiled_v12000 gw_encode_fn()
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

bool dcn_gw_tc_supported(int Cout, int Cin, int KK, int P, long long rows, bool partial_ok)
{
    // partial_ok: P only needs the 16-byte row alignment of the gy tensor map (P % 8); the last 64-pixel k-block of a sample then
    // hangs over its end, where TMA zero-fills the gy operand
    return (partial_ok ? P % 8 == 0 : P % 64 == 0) && Cout % 8 == 0 && (Cin * KK) % 64 == 0 && rows < (1ll << 31) && gw_encode_fn() != nullptr;
}

// halves needed for the gy pairs (the column pairs are written by the scatter kernel into the caller's buffers)
size_t dcn_gw_tc_gy_halves(int B, int Cout, int P) { return 2 * (size_t)B * Cout * P; }

// gw [Cout][Kp] += gy[b0 .. b0+nb) col;  gy_pairs: [hi | lo] halves of the WHOLE batch's gy (split once by dcn_gw_tc_split_gy),
// col_hi / col_lo: [nb*P][Kp] halves of this chunk
int dcn_gw_tc_split_gy(const float *gy, void *gy_pairs, int B, int Cout, int P, cudaStream_t st)
{
    const long long n4 = (long long)B * Cout * P / 4;
    uint2 *hi = reinterpret_cast<uint2 *>(gy_pairs);
    uint32_t *absmax = reinterpret_cast<uint32_t *>(hi + 2 * n4);           // first word behind the pairs (dcn_bwd_cl_gw_floats)
    SIDE_CUDA(cudaMemsetAsync(absmax, 0, sizeof(uint32_t), st));
    const unsigned grid = (unsigned)std::min<long long>((n4 + 255) / 256, 148 * 16);
    gw_absmax_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(gy), absmax, n4);
    SIDE_LAUNCH_CHECK("gw_absmax_kernel");
    gw_split_gy_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(gy), hi, hi + n4, n4, absmax);
    SIDE_LAUNCH_CHECK("gw_split_gy_kernel");
    return SIDE_OK;
}

// Fused convolution weight gradient: gw [Cout][taps * Cq] += gy (*) x, x = the fp16 pairs [N, D, H, W, Cp] of the convolution input read
// through shifted (and, for stride 2, strided) TMA boxes; gy pairs [N, Cout, Do, Ho, Wo].  Returns SIDE_ERR_UNSUPPORTED when the
// output map does not tile into 64-voxel boxes (the caller then builds the patch matrix).
int dcn_gw_tc_run_fused(const void *gy_pairs, const void *x_hi, const void *x_lo, float *gw, int N, int Cout, int Cp, int D, int H, int W,
                        int Do, int Ho, int Wo, int kd, int kh, int kw, int stride, cudaStream_t st)
{
    PFN_cuTensorMapEncodeTiled_v12000 enc = gw_encode_fn();
    SIDE_REQUIRE(enc != nullptr, "dcn_gw_tc: cuTensorMapEncodeTiled is not available from the driver");
    int bw, bh, bd;
    if (!dcn_gw_box(Do, Ho, Wo, bw, bh, bd) || Cp % 8 != 0) {
        set_error("dcn_gw_tc_run_fused: output map %dx%dx%d does not tile into 64-voxel boxes", Do, Ho, Wo);
        return SIDE_ERR_UNSUPPORTED;
    }
    const long long P = (long long)Do * Ho * Wo;
    const int Cq = (Cp + 63) / 64 * 64, taps = kd * kh * kw, Kp = taps * Cq;
    const __half *gh = reinterpret_cast<const __half *>(gy_pairs), *gl = gh + (size_t)N * Cout * P;
    CUtensorMap tm[4];
    {   // gy pairs, box-major rows: [p' (inner), o, n], box {64, 128, 1}; rows past Cout are zero-filled
        cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)Cout, (cuuint64_t)N};
        cuuint64_t strides[2] = {(cuuint64_t)P * 2, (cuuint64_t)Cout * P * 2};
        cuuint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
        for (int i = 0; i < 2; ++i) {
            CUresult r = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half *>(i ? gl : gh), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("dcn_gw_tc: cuTensorMapEncodeTiled(gy, fused) failed (CUresult %d)", (int)r); return SIDE_ERR_UNSUPPORTED; }
        }
    }
    {   // input pairs [C (inner), W, H, D, N], box {64 channels, bw, bh, bd, 1}, traversal stride = the convolution stride on W, H
        cuuint64_t dims[5] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t strides[4] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2, (cuuint64_t)D * H * W * Cp * 2};
        cuuint32_t box[5] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
        cuuint32_t es[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1, 1};
        // with a traversal stride the box extent is given in INPUT elements spanned (b * stride, as conv_tc.cu does): it still
        // delivers bw x bh positions
        box[1] = (cuuint32_t)(bw * stride); box[2] = (cuuint32_t)(bh * stride);
        for (int i = 0; i < 2; ++i) {
            CUresult r = enc(&tm[2 + i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void *>(i ? x_lo : x_hi), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("dcn_gw_tc: cuTensorMapEncodeTiled(x, fused) failed (CUresult %d)", (int)r); return SIDE_ERR_UNSUPPORTED; }
        }
    }
    GwParams p{};
    p.gw = gw; p.Cout = Cout; p.Kp = Kp; p.P = (int)P; p.nb = N;
    p.gy_absmax = reinterpret_cast<const uint32_t *>(gl + (size_t)N * Cout * P);
    p.n_ntiles = (Kp + 255) / 256; p.n_mtiles = (Cout + 127) / 128;
    p.kb_total = N * (int)(P / 64);
    p.fused = 1; p.bw = bw; p.bh = bh; p.bd = bd; p.nbw = Wo / bw; p.nbh = Ho / bh; p.Cq = Cq; p.stride = stride;
    p.kd = kd; p.kh = kh; p.kw = kw;
    const int tiles = p.n_ntiles * p.n_mtiles;
    p.ksplits = std::max(1, std::min(p.kb_total, (296 + tiles - 1) / tiles));
    const size_t smem = (size_t)kGwStages * kGwStage + 1024;
    int rc;
    if ((rc = set_smem_attr((const void *)dcn_gw_tc_kernel, smem))) return rc;
    dcn_gw_tc_kernel<<<(unsigned)(tiles * p.ksplits), kGwThreads, smem, st>>>(tm[0], tm[1], tm[2], tm[3], p);
    SIDE_LAUNCH_CHECK("dcn_gw_tc_kernel<fused conv>");
    return SIDE_OK;
}

int dcn_gw_tc_run(const void *gy_pairs, const void *col_hi, const void *col_lo, float *gw, int B, int b0, int nb, int Cout, int Kp,
                  int P, cudaStream_t st)
{
    PFN_cuTensorMapEncodeTiled_v12000 enc = gw_encode_fn();
    SIDE_REQUIRE(enc != nullptr, "dcn_gw_tc: cuTensorMapEncodeTiled is not available from the driver");
    const __half *gh = reinterpret_cast<const __half *>(gy_pairs), *gl = gh + (size_t)B * Cout * P;
    CUtensorMap tm[4];
    {   // gy pairs of samples b0 .. b0+nb: [p (inner), o, b], box {64, 128, 1}; rows past Cout are zero-filled
        cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)Cout, (cuuint64_t)nb};
        cuuint64_t strides[2] = {(cuuint64_t)P * 2, (cuuint64_t)Cout * P * 2};
        cuuint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
        for (int i = 0; i < 2; ++i) {
            const __half *base = (i ? gl : gh) + (size_t)b0 * Cout * P;
            CUresult r = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half *>(base), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("dcn_gw_tc: cuTensorMapEncodeTiled(gy) failed (CUresult %d)", (int)r); return SIDE_ERR_CUDA; }
        }
    }
    {   // column pairs: [k' (inner), pixel row], box {64 k', 64 pixels}
        cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)nb * P};
        cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
        cuuint32_t box[2] = {64, 64}, es[2] = {1, 1};
        for (int i = 0; i < 2; ++i) {
            CUresult r = enc(&tm[2 + i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(i ? col_lo : col_hi), dims, strides, box,
                             es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("dcn_gw_tc: cuTensorMapEncodeTiled(col) failed (CUresult %d)", (int)r); return SIDE_ERR_CUDA; }
        }
    }
    GwParams p{};
    p.gw = gw; p.Cout = Cout; p.Kp = Kp; p.P = P; p.nb = nb;
    p.gy_absmax = reinterpret_cast<const uint32_t *>(gl + (size_t)B * Cout * P);
    p.n_ntiles = (Kp + 255) / 256; p.n_mtiles = (Cout + 127) / 128;
    p.kb_total = nb * ((P + 63) / 64);
    const int tiles = p.n_ntiles * p.n_mtiles;
    p.ksplits = std::max(1, std::min(p.kb_total, (296 + tiles - 1) / tiles));
    const size_t smem = (size_t)kGwStages * kGwStage + 1024;
    int rc;
    if ((rc = set_smem_attr((const void *)dcn_gw_tc_kernel, smem))) return rc;
    dcn_gw_tc_kernel<<<(unsigned)(tiles * p.ksplits), kGwThreads, smem, st>>>(tm[0], tm[1], tm[2], tm[3], p);
    SIDE_LAUNCH_CHECK("dcn_gw_tc_kernel");
    return SIDE_OK;
}

}  // namespace side
