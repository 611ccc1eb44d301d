// dcn_fwd_tc.cu -- DCNv2 forward on the 5th-gen tensor cores: tcgen05.mma (kind::tf32) with the accumulator in TMEM.
//
// The contraction y[pixel, o] = sum_{tap, c} col[pixel, tap, c] * W[o, c, tap] is a dense GEMM
// (M = B*Ho*Wo pixels, N = Cout, K = kh*kw*Cin) whose A operand does not exist in memory: it is the modulated
// bilinear gather of the reference's im2col kernel (dcn_v2_im2col_cuda.cu:125-195).  Warp-specialised CTA, one
// 128-pixel tile per CTA:
//   warps 0-7  producers : thread = (pixel row, half of the 32-channel K block).  Computes the tap geometry once
//                          per tap, gathers 16 channels (coalesced along W across the warp), multiplies by the mask
//                          and writes the values STRAIGHT INTO THE UMMA OPERAND LAYOUT in shared memory
//                          (K-major, 128-byte swizzle: one 128-byte row = 32 tf32 of one pixel, 16-byte chunk index
//                          XOR (row & 7)), then fence.proxy.async + mbarrier arrive.  No `columns` tensor in HBM.
//   warp 8     MMA issuer: one elected thread issues tcgen05.mma cta_group::1, M=128, N=Cout, K=8 per instruction,
//                          D in TMEM (128 lanes x Cout fp32 columns); tcgen05.commit releases the smem stage.
//   warp 9     weight loader: the weights were re-laid-out once into per-K-block tiles that ARE the shared-memory
//                          image (swizzle included), so each stage is ONE TMA bulk copy (cp.async.bulk, UBLKCP)
//                          completing on the stage's mbarrier.
//   epilogue   (warps 0-7 again) tcgen05.ld 32x32b -> bias / folded BatchNorm / ReLU -> coalesced NCHW stores
//              (a warp's 32 lanes are 32 consecutive pixels of one output channel).
//
// Precision.  tcgen05 has no fp32 MMA.  SIDE_DCN_PREC_3XTF32 splits both operands into hi = top 19 bits and
// lo = x - hi (exact in fp32) and issues hi*hi + lo*hi + hi*lo into the same fp32 accumulator: the dropped lo*lo
// term is 2^-20 relative, so the result meets the reference's fp32 tolerance (<= 1e-4 relative; measured ~1e-6).
// SIDE_DCN_PREC_TF32 is the single-pass opt-in (~1e-3).
#include <algorithm>
#include "tc_common.cuh"

namespace side {

constexpr int kTcBM = 128;            // UMMA M (pixels per tile)
constexpr int kTcProducerThreads = 512;
constexpr int kTcThreads = kTcProducerThreads + 64;
constexpr int kTcMaxStages = 8;
constexpr int kTcMaxTapGroup = 9;     // taps whose sample geometry is computed together (one exposed latency per group)
constexpr uint32_t kATileBytes = kTcBM * 128;   // 16 KB

// w [Cout, Cin, KK] -> per-K-block tiles: wp[kb][part][Cout x 32] in the swizzled shared-memory image.
// kb = tap * (Cin/32) + cb;  part 0 = hi (or the full value when !split), part 1 = lo.
__global__ void dcn_tc_weight_prep_kernel(const float *__restrict__ w, float *__restrict__ wp, int Cout, int Cin, int KK,
                                          int split)
{
    const long long total = (long long)KK * Cin * Cout;
    const int ncb = Cin / kTcBK;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i % kTcBK);
        const int n = (int)((i / kTcBK) % Cout);
        const int kb = (int)(i / ((long long)kTcBK * Cout));
        const int tap = kb / ncb, cb = kb - tap * ncb;
        const float v = __ldg(w + ((size_t)n * Cin + cb * kTcBK + kk) * KK + tap);
        const size_t tile = (size_t)kb * (split ? 2 : 1) * Cout * kTcBK;
        const uint32_t off = (sw128(n, kk >> 2) >> 2) + (kk & 3);
        if (split) {
            const float hi = tf32_hi(v);
            wp[tile + off] = hi;
            wp[tile + (size_t)Cout * kTcBK + off] = v - hi;
        } else {
            wp[tile + off] = v;
        }
    }
}

// fp16-pair variant (SIDE_DCN_PREC_3XFP16): k-block = 64 channels (64 halfs = one 128-byte swizzle row), part 0 = hi = fp16(w),
// part 1 = lo' = fp16((w - hi) * 2^11) (tc_common.cuh).  max |w| goes to the range-guard slot.
__global__ void dcn_tc_weight_prep_f16_kernel(const float *__restrict__ w, __half *__restrict__ wp, int Cout, int Cin, int KK,
                                              uint32_t *rs)
{
    const long long total = (long long)KK * Cin * Cout;
    const int ncb = Cin / 64;
    float amax = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i % 64);
        const int n = (int)((i / 64) % Cout);
        const int kb = (int)(i / ((long long)64 * Cout));
        const int tap = kb / ncb, cb = kb - tap * ncb;
        const float v = __ldg(w + ((size_t)n * Cin + cb * 64 + kk) * KK + tap);
        amax = fmaxf(amax, fabsf(v));
        const size_t tile = (size_t)kb * 2 * Cout * 64;                       // halfs
        const uint32_t off = (sw128(n, kk >> 3) >> 1) + (kk & 7);
        __half h, l;
        f16_split(v, h, l);
        wp[tile + off] = h;
        wp[tile + (size_t)Cout * 64 + off] = l;
    }
    if (rs) range_commit(rs, amax);
}

// sample geometry of one (pixel, tap), ready for the channels-last gather: element offsets into the NHWC copy of x
// (batch folded in) and bilinear weights pre-multiplied by the modulation mask
struct TapRec {
    uint32_t o1, o2, o3, o4;     // BYTE offsets of the four corners' pixels (channel 0) in the channels-last copy
    float w1, w2, w3, w4;
};

// Tile geometry: 128 pixels = th x tw.  8 x 16 tiles keep the footprint of the 9 deformed taps compact (for small
// offsets ~10 x 18 source pixels x Cin channels), so that with only 2-3 operand stages in shared memory the rest of
// the 228 KB stays L1 and most of the 36 corner reads per pixel hit it; maps that do not tile into 8 x 16 use 1 x 128
// over the flattened (b, y, x) index.
struct TcTiling {
    int th, tw;          // tile extent (th * tw == 128); th == 1: flattened
    int tiles_x, tiles_y;
};

// Split-K over the taps for small maps (12x40 / 24x80 layers at small batch: 12 .. 90 pixel tiles on 148 SMs): ksplit CTAs
// share a pixel tile, each contracts a contiguous range of taps and writes its raw sums to `partial`; the CTA that takes the
// tile's last ticket adds the partials in split order (deterministic) and runs the bias / BatchNorm / ReLU epilogue.
struct TcSplitK {
    int ksplit;                // 1 = off
    float *partial;            // [ksplit][B][Cout][P]
    unsigned int *tickets;     // [tiles], zeroed before the launch
};

// F16: operands are fp16 (hi, lo' = lo * 2^11) pairs, kind::f16 -- a k-block is 64 channels (the same 128-byte rows), so the
// tile takes half the MMA instructions, half the operand bytes and half the barrier round trips of the tf32 pairs.
template <bool SPLIT, bool F16>
__global__ void __launch_bounds__(kTcThreads, 1) dcn_fwd_tc_kernel(DcnFwdArgs a, const float *__restrict__ wp,
                                                                  const float *__restrict__ xt, int stages, uint32_t idesc,
                                                                  uint32_t tmem_cols, TcTiling tl, int tap_group, TcSplitK sk,
                                                                  uint32_t *rs)
{
    static_assert(SPLIT || !F16, "fp16 operands are always hi/lo pairs");
    constexpr int kKbCh = F16 ? 64 : kTcBK;            // channels per k-block
    constexpr int kSub = F16 ? 2 : 1;                  // 32-channel producer passes per k-block
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[kTcMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kTcMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ unsigned int s_ticket;

    const DcnShape &s = a.s;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler as well
    const int N = s.Cout;
    const uint32_t b_part_bytes = (uint32_t)N * 128u;
    const uint32_t a_bytes = SPLIT ? 2 * kATileBytes : kATileBytes;
    const uint32_t b_bytes = SPLIT ? 2 * b_part_bytes : b_part_bytes;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    // SWIZZLE_128B tiles must start on a 1024-byte boundary of the SHARED address space
    unsigned char *tiles = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    TapRec *tapbuf = reinterpret_cast<TapRec *>(tiles + (size_t)stages * stage_bytes);      // [tap_group][128]

    const int ncb = s.Cin / kKbCh;
    const int tile = (int)blockIdx.x / sk.ksplit, split = (int)blockIdx.x - tile * sk.ksplit;
    const int tap_lo = split * s.KK / sk.ksplit, tap_hi = (split + 1) * s.KK / sk.ksplit;      // this CTA's taps
    const int nkb = (tap_hi - tap_lo) * ncb;

    if (tid == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], kTcProducerThreads / 32 + 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        mbar_fence_init();
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_smem;

    if (tid < kTcProducerThreads) {
        // ================= producers: modulated bilinear gather -> swizzled A tile =================
        int tile_b = 0, tile_y0 = 0, tile_x0 = 0;
        if (tl.th != 1) {
            const int per_img = tl.tiles_x * tl.tiles_y;
            tile_b = tile / per_img;
            const int t = tile - tile_b * per_img, ty = t / tl.tiles_x;
            tile_y0 = ty * 8; tile_x0 = (t - ty * tl.tiles_x) * 16;
        }
        // pixel of tile row m
        auto pixel_of = [&](int m, int &b, int &ho, int &wo) -> bool {
            if (tl.th == 1) {
                const long long gp = (long long)tile * kTcBM + m;
                if (gp >= (long long)s.B * s.P) return false;
                b = (int)(gp / s.P);
                const int p = (int)(gp - (long long)b * s.P);
                ho = p / s.Wo; wo = p - ho * s.Wo;
                return true;
            }
            b = tile_b;
            ho = tile_y0 + (m >> 4); wo = tile_x0 + (m & 15);       // 8 x 16 tiles
            return ho < s.Ho && wo < s.Wo;
        };
        // gather: thread = (c4 = tid & 7 -> 4 channels, rows (tid >> 3) and (tid >> 3) + 64): a warp instruction reads 4 pixels x
        // 128 contiguous bytes of the channels-last copy
        const int c4 = tid & 7, r0 = tid >> 3;
        float4 raw[8];
        auto geometry = [&](int tap0) {         // sample geometry of taps tap0 .. tap0 + tap_group - 1 for all 128 pixels
            const int ntap = min(tap_group, tap_hi - tap0);
            for (int i = tid; i < ntap * kTcBM; i += kTcProducerThreads) {
                const int tt = i / kTcBM, m = i - tt * kTcBM;
                int b, ho, wo;
                TapRec tr;
                if (pixel_of(m, b, ho, wo)) {
                    const DcnTap g = dcn_tap(s, a.offset, a.mask, b, 0, tap0 + tt, ho, wo);
                    const int base = b * s.H * s.W;
                    const uint32_t pix = (uint32_t)s.Cin * 4u;
                    tr.o1 = (uint32_t)(base + g.o1) * pix; tr.o2 = (uint32_t)(base + g.o2) * pix;
                    tr.o3 = (uint32_t)(base + g.o3) * pix; tr.o4 = (uint32_t)(base + g.o4) * pix;
                    tr.w1 = g.w1 * g.m; tr.w2 = g.w2 * g.m; tr.w3 = g.w3 * g.m; tr.w4 = g.w4 * g.m;
                } else {
                    tr.o1 = tr.o2 = tr.o3 = tr.o4 = 0u;
                    tr.w1 = tr.w2 = tr.w3 = tr.w4 = 0.f;
                }
                tapbuf[tt * kTcBM + m] = tr;
            }
        };
        // issue the 8 corner loads of one 32-channel pass (tap slot `slot` of the table, 32-channel block cb)
        auto prefetch = [&](int slot, int cb) {
            const TapRec *tb = tapbuf + slot * kTcBM;
            const char *xc = reinterpret_cast<const char *>(xt + cb * kTcBK + 4 * c4);   // one 64-bit base per pass,
#pragma unroll
            for (int i = 0; i < 2; ++i) {                                                // unsigned 32-bit byte offsets per corner
                const uint4 o = *reinterpret_cast<const uint4 *>(&tb[r0 + 64 * i]);
                raw[4 * i + 0] = __ldg(reinterpret_cast<const float4 *>(xc + o.x));
                raw[4 * i + 1] = __ldg(reinterpret_cast<const float4 *>(xc + o.y));
                raw[4 * i + 2] = __ldg(reinterpret_cast<const float4 *>(xc + o.z));
                raw[4 * i + 3] = __ldg(reinterpret_cast<const float4 *>(xc + o.w));
            }
        };
        geometry(tap_lo);
        asm volatile("bar.sync 1, %0;" ::"n"(kTcProducerThreads) : "memory");
        prefetch(0, 0);
        // all loop state is carried incrementally: no integer division in the k loop.  cb counts 32-channel passes of the tap
        const int npass = ncb * kSub;
        float amax = 0.f;
        int st = 0, tp = tap_lo, cb = 0, slot = 0;
        uint32_t eph = 1;                        // parity to wait for on empty_bar[st] (first pass falls through)
        // swizzled byte offset of this thread's 8 (fp16) / 16 (tf32) bytes inside the A tile, row r0: row r0 + 64 is 8 KB further
        // (same row & 7), the second 32-channel half of an fp16 k-block flips chunk bit 2 = byte 64.  Computed once -- inside the
        // loop the compiler rebuilt it from threadIdx every pass (112 registers per thread at 576 threads).
        const uint32_t off0 = F16 ? sw128(r0, c4 >> 1) + ((c4 & 1) << 3) : sw128(r0, c4);
        for (int kb = 0; kb < nkb * kSub; ++kb) {
            const TapRec *tb = tapbuf + slot * kTcBM;
            float4 v[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float4 w = *reinterpret_cast<const float4 *>(&tb[r0 + 64 * i].w1);
                const float4 a1 = raw[4 * i], a2 = raw[4 * i + 1], a3 = raw[4 * i + 2], a4 = raw[4 * i + 3];
                v[i].x = fmaf(w.w, a4.x, fmaf(w.z, a3.x, fmaf(w.y, a2.x, w.x * a1.x)));
                v[i].y = fmaf(w.w, a4.y, fmaf(w.z, a3.y, fmaf(w.y, a2.y, w.x * a1.y)));
                v[i].z = fmaf(w.w, a4.z, fmaf(w.z, a3.z, fmaf(w.y, a2.z, w.x * a1.z)));
                v[i].w = fmaf(w.w, a4.w, fmaf(w.z, a3.w, fmaf(w.y, a2.w, w.x * a1.w)));
            }
            // `raw` is dead from here on: the next pass's loads below go into the same registers.  Without this fence the compiler
            // hoists them above the FMAs into fresh registers and pays 32 register moves per pass at the end of the loop.
            asm volatile("" ::: "memory");
            // advance (tap, channel block) to the next k-block and start its loads while this one is stored
            int ncbi = cb + 1, ntp = tp, nslot = slot;
            if (ncbi == npass) {
                ncbi = 0; ++ntp;
                if (++nslot == tap_group) nslot = 0;
            }
            if (kb + 1 < nkb * kSub) {
                if (ntp != tp && nslot == 0) {              // the next k-block starts a new tap group: recompute the table
                    asm volatile("bar.sync 1, %0;" ::"n"(kTcProducerThreads) : "memory");   // everyone is done with the old one
                    geometry(ntp);
                    asm volatile("bar.sync 1, %0;" ::"n"(kTcProducerThreads) : "memory");
                }
                prefetch(nslot, ncbi);
            }
            const int half = F16 ? (cb & 1) : 0;                 // which 32 channels of the 64-channel k-block
            if (!F16 || half == 0) mbar_wait(&empty_bar[st], eph);
            unsigned char *sa = tiles + (size_t)st * stage_bytes;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (F16) {
                    // 4 channels = 8 bytes of the row's 128: chunk 4 half + c4 / 2, upper or lower 8 bytes
                    const uint32_t off = (off0 ^ ((uint32_t)half << 6)) + (uint32_t)i * 8192u;
                    uint2 hi, lo;
                    f16_split2(v[i].x, v[i].y, hi.x, lo.x);
                    f16_split2(v[i].z, v[i].w, hi.y, lo.y);
                    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
                    *reinterpret_cast<uint2 *>(sa + off) = hi;
                    *reinterpret_cast<uint2 *>(sa + kATileBytes + off) = lo;
                    continue;
                }
                const uint32_t off = off0 + (uint32_t)i * 8192u;
                if (SPLIT) {
                    float4 hi, lo;
                    hi.x = tf32_hi(v[i].x); lo.x = v[i].x - hi.x;
                    hi.y = tf32_hi(v[i].y); lo.y = v[i].y - hi.y;
                    hi.z = tf32_hi(v[i].z); lo.z = v[i].z - hi.z;
                    hi.w = tf32_hi(v[i].w); lo.w = v[i].w - hi.w;
                    *reinterpret_cast<float4 *>(sa + off) = hi;
                    *reinterpret_cast<float4 *>(sa + kATileBytes + off) = lo;
                } else {
                    *reinterpret_cast<float4 *>(sa + off) = v[i];
                }
            }
            if (!F16 || half == 1) {
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[st]);   // one arrival per producer warp
                if (++st == stages) { st = 0; eph ^= 1u; }
            }
            cb = ncbi; tp = ntp; slot = nslot;
        }
        if (F16 && rs) range_commit(rs, amax);

        // ================= epilogue: TMEM -> registers -> NCHW =================
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const int lg = warp & 3, cq = warp >> 2;             // TMEM lane group of this warp, column quarter
        const bool affine = s.flags & SIDE_DCN_FUSE_AFFINE, relu = s.flags & SIDE_DCN_FUSE_RELU;
        int b, ho, wo;
        const bool pix_ok = pixel_of(lg * 32 + lane, b, ho, wo);
        float *yp = a.y + (pix_ok ? (size_t)b * s.Cout * s.P + (size_t)ho * s.Wo + wo : 0);
        const int cbeg = cq * (N / 4), cend = cbeg + N / 4;
        const float acc_fix = tc_acc_fix(nkb * (kTcBK / 8));            // truncating accumulation, tc_common.cuh
        const size_t pix_off = pix_ok ? (size_t)b * s.Cout * s.P + (size_t)ho * s.Wo + wo : 0;
        float *pp = sk.ksplit > 1 ? sk.partial + (size_t)split * s.B * s.Cout * s.P + pix_off : nullptr;
        for (int c = cbeg; c < cend; c += 8) {
            float acc[8];
            tc_ld8(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)c, acc);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= acc_fix;
            if (SPLIT) {                      // cross-term accumulator (see the MMA issuer)
                float accx[8];
                tc_ld8(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(N + c), accx);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = F16 ? fmaf(accx[j], kF16LoInv, acc[j]) : acc[j] + accx[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = c + j;
                if (n >= cend) break;                          // Cout = 16: a column quarter is 4 wide, the load 8
                if (pp) {                                      // split-K: raw partial sum, finished by the tile's last CTA
                    if (pix_ok) __stcg(pp + (size_t)n * s.P, acc[j]);
                    continue;
                }
                float o = acc[j] + (a.bias ? __ldg(a.bias + n) : 0.f);
                if (affine) o = fmaf(o, __ldg(a.scale + n), __ldg(a.shift + n));
                if (relu) o = fmaxf(o, 0.f);
                if (pix_ok) yp[(size_t)n * s.P] = o;
            }
        }
        if (sk.ksplit > 1) {
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(kTcProducerThreads) : "memory");
            if (tid == 0) s_ticket = atomicAdd(sk.tickets + tile, 1u);
            asm volatile("bar.sync 1, %0;" ::"n"(kTcProducerThreads) : "memory");
            if (s_ticket == (unsigned)sk.ksplit - 1u) {
                __threadfence();
                const size_t split_stride = (size_t)s.B * s.Cout * s.P;
                // 4 channels x up to 9 partials = 36 independent loads in flight per thread (a loop with a run-time trip count
                // serialised them: ~600 cycles of L2 latency per load made this tail longer than the contraction itself);
                // the partials are still added in split order, so the result does not depend on which CTA finishes last
                for (int n0 = cbeg; n0 < cend; n0 += 4) {
                    float v[4][9];
#pragma unroll
                    for (int k = 0; k < 9; ++k)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            v[j][k] = (pix_ok && k < sk.ksplit && n0 + j < cend)
                                          ? __ldcg(sk.partial + (size_t)k * split_stride + pix_off + (size_t)(n0 + j) * s.P) : 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = n0 + j;
                        if (n >= cend) break;
                        float o = 0.f;
#pragma unroll
                        for (int k = 0; k < 9; ++k) o += v[j][k];
                        o += a.bias ? __ldg(a.bias + n) : 0.f;
                        if (affine) o = fmaf(o, __ldg(a.scale + n), __ldg(a.shift + n));
                        if (relu) o = fmaxf(o, 0.f);
                        if (pix_ok) yp[(size_t)n * s.P] = o;
                    }
                }
            }
        }
    } else if (warp == 16) {
        // ================= MMA issuer =================
        // all 32 lanes run the loop (uniform operands), the elected lane issues -- see tc_elect_one()
        {
            const bool leader = tc_elect_one();
            const uint32_t idesc2 = tc_idesc<F16>(kTcBM, 2 * N <= 256 ? 2 * N : N);
            const uint64_t desc0 = tc_smem_desc(smem_u32(tiles));
            int st = 0;
            uint32_t fph = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[st], fph);
                tc_fence_after();
                const uint64_t sa = tc_desc_add(desc0, (uint32_t)st * stage_bytes);
                const uint64_t sb = tc_desc_add(sa, a_bytes);
                if (leader) {
#pragma unroll
                    for (int k = 0; k < kTcBK / 8; ++k) {
                        const uint64_t a_hi = tc_desc_add(sa, k * 32), b_hi = tc_desc_add(sb, k * 32);
                        if (SPLIT && 2 * N <= 256) {
                            // hi*hi and hi*lo go out as ONE instruction of width 2N against the adjacent [B_hi | B_lo] tiles (main ->
                            // columns 0..N, cross -> N..2N), lo*hi then accumulates onto the cross columns: 2 instructions per k-step
                            const uint64_t a_lo = tc_desc_add(sa, kATileBytes + k * 32);
                            tc_mma<F16>(tmem_d, a_hi, b_hi, idesc2, (kb | k) != 0 ? 1u : 0u);
                            tc_mma<F16>(tmem_d + (uint32_t)N, a_lo, b_hi, idesc, 1u);
                            continue;
                        }
                        tc_mma<F16>(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        if (SPLIT) {
                            // The two small cross terms go to a SECOND accumulator (columns N..2N): the tensor core adds
                            // into fp32 with truncation, ~0.5 ulp of the accumulator per MMA, so keeping them out of the
                            // main accumulator cuts that systematic bias 3x; the epilogue adds the two.
                            const uint64_t a_lo = tc_desc_add(sa, kATileBytes + k * 32);
                            const uint64_t b_lo = tc_desc_add(sb, b_part_bytes + k * 32);
                            tc_mma<F16>(tmem_d + (uint32_t)N, a_lo, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                            tc_mma<F16>(tmem_d + (uint32_t)N, a_hi, b_lo, idesc, 1u);
                        }
                    }
                    tc_commit(&empty_bar[st]);     // smem stage reusable once these MMAs have read it
                }
                if (++st == stages) { st = 0; fph ^= 1u; }
            }
            if (leader) tc_commit(&tmem_full_bar);         // accumulator complete
        }
        __syncwarp();
    } else {
        // ================= weight loader (TMA bulk copies) =================
        {
            const bool leader = tc_elect_one();
            int st = 0;
            uint32_t eph = 1;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&empty_bar[st], eph);
                if (leader) {
                    mbar_expect_tx(&full_bar[st], b_bytes);
                    bulk_g2s(tiles + (size_t)st * stage_bytes + a_bytes, wp + (size_t)(tap_lo * ncb + kb) * (b_bytes / 4), b_bytes, &full_bar[st]);
                }
                if (++st == stages) { st = 0; eph ^= 1u; }
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}

int launch_tc_weight_prep(const float *w, float *wp, int Cout, int Cin, int KK, int split, cudaStream_t st)
{
    const long long nW = (long long)Cout * Cin * KK;
    dcn_tc_weight_prep_kernel<<<(unsigned)std::min<long long>(1184, (nW + 255) / 256), 256, 0, st>>>(w, wp, Cout, Cin, KK, split);
    SIDE_LAUNCH_CHECK("dcn_tc_weight_prep_kernel");
    return SIDE_OK;
}

bool dcn_fwd_tc_supported(int Cin, int Cout, int dg)
{
    return dg == 1 && Cin % kTcBK == 0 && Cout % 16 == 0 && Cout >= 16 && Cout <= 256;
}

// fp16 pairs need whole 64-channel k-blocks; other channel counts run the request as 3xTF32
static int tc_precision(int Cin, int flags)
{
    const int prec = flags & SIDE_DCN_PREC_MASK;
    return (prec == SIDE_DCN_PREC_3XFP16 && Cin % 64 != 0) ? SIDE_DCN_PREC_3XTF32 : prec;
}
static size_t tc_weight_bytes(int Cin, int Cout, int KK, int flags)
{
    // hi + lo tf32 words (3xTF32), hi + lo fp16 halves (3xFP16: half of that), or the plain fp32 words (TF32 single pass)
    const bool split = tc_precision(Cin, flags) != SIDE_DCN_PREC_TF32;
    return (sizeof(float) * (split ? 2 : 1) * (size_t)Cin * Cout * KK + 255) & ~(size_t)255;
}

// pixel tiles of a launch, and how many CTAs should share one (split-K over the taps) to fill the GPU
static long long tc_tiles(int B, int Ho, int Wo)
{
    if (Ho % 8 == 0 && Wo % 16 == 0) return (long long)B * (Ho / 8) * (Wo / 16);
    return ((long long)B * Ho * Wo + kTcBM - 1) / kTcBM;
}
static int tc_ksplit(long long tiles, int KK)
{
    if (tiles >= 100 || KK % 3 != 0) return 1;
    if (tiles * 3 >= 100 || KK % 9 != 0) return 3;
    return 9;
}
static size_t tc_split_bytes(int B, int Cout, int Ho, int Wo, int KK)
{
    const long long tiles = tc_tiles(B, Ho, Wo);
    const int ks = tc_ksplit(tiles, KK);
    if (ks == 1) return 0;
    return sizeof(float) * (size_t)ks * B * Cout * Ho * Wo + ((sizeof(unsigned int) * (size_t)tiles + 255) & ~(size_t)255);
}

// workspace = weight tiles + the channels-last copy of x (+ split-K partial sums and tickets for small maps; sized for
// same-size output, the upper bound for stride >= 1)
size_t dcn_fwd_tc_ws_bytes(int B, int Cin, int H, int W, int Cout, int KK, int flags)
{
    return tc_weight_bytes(Cin, Cout, KK, flags) + ((sizeof(float) * (size_t)B * Cin * H * W + 255) & ~(size_t)255) +
           tc_split_bytes(B, Cout, H, W, KK);
}

int dcn_fwd_tc(const DcnFwdArgs &a, const float *w, void *ws, size_t ws_bytes, cudaStream_t st, const float *x_nhwc)
{
    const DcnShape &s = a.s;
    const int prec = tc_precision(s.Cin, s.flags);
    const bool f16 = prec == SIDE_DCN_PREC_3XFP16, split = prec != SIDE_DCN_PREC_TF32;
    if (!dcn_fwd_tc_supported(s.Cin, s.Cout, s.dg)) {
        set_error("side_dcn_fwd: tcgen05 path needs dg == 1, Cin %% 32 == 0, Cout %% 16 == 0, 16 <= Cout <= 256");
        return SIDE_ERR_UNSUPPORTED;
    }
    const size_t base_bytes = tc_weight_bytes(s.Cin, s.Cout, s.KK, s.flags) + ((sizeof(float) * (size_t)s.B * s.Cin * s.H * s.W + 255) & ~(size_t)255);
    if (ws_bytes < base_bytes) {
        set_error("side_dcn_fwd: workspace too small for the tcgen05 weight tiles + NHWC copy");
        return SIDE_ERR_WORKSPACE;
    }
    if ((long long)s.B * s.H * s.W * s.Cin * 4 >= (1ll << 32)) {
        set_error("side_dcn_fwd: input too large for 32-bit gather byte offsets");
        return SIDE_ERR_UNSUPPORTED;
    }
    float *wp = reinterpret_cast<float *>(ws);
    float *xt = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(ws) + tc_weight_bytes(s.Cin, s.Cout, s.KK, s.flags));
    int rc = SIDE_OK;
    if (x_nhwc) xt = const_cast<float *>(x_nhwc);           // the caller already holds the channels-last copy
    else if ((rc = launch_nchw_to_nhwc(a.x, xt, s.B, s.Cin, s.H * s.W, st))) return rc;
    uint32_t *rs = f16 ? range_slot_next() : nullptr;
    if (f16) {
        const long long nW = (long long)s.Cout * s.Cin * s.KK;
        dcn_tc_weight_prep_f16_kernel<<<(unsigned)std::min<long long>(1184, (nW + 255) / 256), 256, 0, st>>>(
            w, reinterpret_cast<__half *>(wp), s.Cout, s.Cin, s.KK, rs);
        SIDE_LAUNCH_CHECK("dcn_tc_weight_prep_f16_kernel");
        rs = range_slot_next();                     // the gathered activations report into their own slot
    } else if ((rc = launch_tc_weight_prep(w, wp, s.Cout, s.Cin, s.KK, split ? 1 : 0, st))) return rc;

    const uint32_t stage_bytes = (split ? 2u : 1u) * (kATileBytes + (uint32_t)s.Cout * 128u);
    // operand stages: the gather, not the MMA, bounds this kernel, so 2-3 stages are enough; what is left of the 228 KB
    // stays L1 for the 36 corner reads per pixel.  The tap table holds all taps when it fits next to the stages.
    int stages = std::max(2, std::min((int)((100u * 1024u) / stage_bytes), 3));
    int tap_group = std::min(s.KK, kTcMaxTapGroup);
    while (tap_group > 1 && (size_t)stages * stage_bytes + (size_t)tap_group * kTcBM * sizeof(TapRec) + 1024 > 220 * 1024) --tap_group;
    const size_t smem = (size_t)stages * stage_bytes + (size_t)tap_group * kTcBM * sizeof(TapRec) + 1024;
    if (smem > 226 * 1024) {
        set_error("side_dcn_fwd: tcgen05 path does not fit shared memory for Cout=%d", s.Cout);
        return SIDE_ERR_UNSUPPORTED;
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < (split ? 2 : 1) * s.Cout) tmem_cols <<= 1;
    const uint32_t idesc = f16 ? tc_idesc_f16(kTcBM, s.Cout) : tc_idesc_tf32(kTcBM, s.Cout);
    TcTiling tl;
    unsigned grid;
    if (s.Ho % 8 == 0 && s.Wo % 16 == 0) {
        tl.th = 8; tl.tw = 16; tl.tiles_x = s.Wo / 16; tl.tiles_y = s.Ho / 8;
        grid = (unsigned)((long long)s.B * tl.tiles_x * tl.tiles_y);
    } else {
        tl.th = 1; tl.tw = 128; tl.tiles_x = tl.tiles_y = 0;
        grid = (unsigned)ceil_div((long long)s.B * s.P, kTcBM);
    }
    // split-K over the taps when the pixel tiles alone leave most SMs idle and the workspace has room for the partial sums
    TcSplitK sk{1, nullptr, nullptr};
    {
        const int ks = tc_ksplit(grid, s.KK);
        const size_t part_bytes = sizeof(float) * (size_t)ks * s.B * s.Cout * s.P;
        const size_t tick_bytes = (sizeof(unsigned int) * (size_t)grid + 255) & ~(size_t)255;
        if (ks > 1 && ws_bytes >= base_bytes + part_bytes + tick_bytes) {
            unsigned char *pb = reinterpret_cast<unsigned char *>(ws) + base_bytes;
            sk.ksplit = ks;
            sk.partial = reinterpret_cast<float *>(pb);
            sk.tickets = reinterpret_cast<unsigned int *>(pb + part_bytes);
            SIDE_CUDA(cudaMemsetAsync(sk.tickets, 0, sizeof(unsigned int) * (size_t)grid, st));
            grid *= (unsigned)ks;
        }
    }
    if (f16) {
        if ((rc = set_smem_attr((const void *)dcn_fwd_tc_kernel<true, true>, smem))) return rc;
        dcn_fwd_tc_kernel<true, true><<<grid, kTcThreads, smem, st>>>(a, wp, xt, stages, idesc, tmem_cols, tl, tap_group, sk, rs);
    } else if (split) {
        if ((rc = set_smem_attr((const void *)dcn_fwd_tc_kernel<true, false>, smem))) return rc;
        dcn_fwd_tc_kernel<true, false><<<grid, kTcThreads, smem, st>>>(a, wp, xt, stages, idesc, tmem_cols, tl, tap_group, sk, rs);
    } else {
        if ((rc = set_smem_attr((const void *)dcn_fwd_tc_kernel<false, false>, smem))) return rc;
        dcn_fwd_tc_kernel<false, false><<<grid, kTcThreads, smem, st>>>(a, wp, xt, stages, idesc, tmem_cols, tl, tap_group, sk, rs);
    }
    SIDE_LAUNCH_CHECK("dcn_fwd_tc_kernel");
    return SIDE_OK;
}

}  // namespace side
