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
constexpr int kTcProducerThreads = 256;
constexpr int kTcThreads = kTcProducerThreads + 64;
constexpr int kTcMaxStages = 8;
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

// sample geometry of one (pixel, tap), ready for the channels-last gather: element offsets into the NHWC copy of x
// (batch folded in) and bilinear weights pre-multiplied by the modulation mask
struct TapRec {
    int o1, o2, o3, o4;
    float w1, w2, w3, w4;
};

template <bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 1) dcn_fwd_tc_kernel(DcnFwdArgs a, const float *__restrict__ wp,
                                                                  const float *__restrict__ xt, int stages, uint32_t idesc,
                                                                  uint32_t tmem_cols)
{
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[kTcMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kTcMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) TapRec tapbuf[2][kTcBM];

    const DcnShape &s = a.s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = s.Cout;
    const uint32_t b_part_bytes = (uint32_t)N * 128u;
    const uint32_t a_bytes = SPLIT ? 2 * kATileBytes : kATileBytes;
    const uint32_t b_bytes = SPLIT ? 2 * b_part_bytes : b_part_bytes;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    // SWIZZLE_128B tiles must start on a 1024-byte boundary of the SHARED address space
    unsigned char *tiles = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);

    const int ncb = s.Cin / kTcBK;
    const int nkb = s.KK * ncb;

    if (tid == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], kTcProducerThreads / 32 + 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        mbar_fence_init();
    }
    if (warp == 8) {
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
        // geometry: thread `row` (< 128) fills tapbuf once per tap; gather: thread = (c4 = tid & 7, rows (tid >> 3) + 32 i):
        // a warp instruction reads 4 pixels x 128 contiguous bytes of the NHWC copy.
        const int row = tid & 127;
        const long long Mtot = (long long)s.B * s.P;
        const long long gp = (long long)blockIdx.x * kTcBM + row;
        const bool pix_ok = gp < Mtot;
        const int b = pix_ok ? (int)(gp / s.P) : 0;
        const int p = pix_ok ? (int)(gp - (long long)b * s.P) : 0;
        const int ho = p / s.Wo, wo = p - ho * s.Wo;
        const int c4 = tid & 7, r0 = tid >> 3;
        int cur_tap = -1;
        float4 raw[16];
        // issue the 16 tap loads of k-block kb (and publish the tap geometry when a new tap starts)
        auto prefetch = [&](int kb) {
            const int tp = kb / ncb, c0 = (kb - tp * ncb) * kTcBK + 4 * c4;
            if (tp != cur_tap) {
                cur_tap = tp;
                if (tid < kTcBM) {
                    TapRec tr;
                    if (pix_ok) {
                        const DcnTap g = dcn_tap(s, a.offset, a.mask, b, 0, tp, ho, wo);
                        const int base = b * s.H * s.W;
                        tr.o1 = (base + g.o1) * s.Cin; tr.o2 = (base + g.o2) * s.Cin;
                        tr.o3 = (base + g.o3) * s.Cin; tr.o4 = (base + g.o4) * s.Cin;
                        tr.w1 = g.w1 * g.m; tr.w2 = g.w2 * g.m; tr.w3 = g.w3 * g.m; tr.w4 = g.w4 * g.m;
                    } else {
                        tr.o1 = tr.o2 = tr.o3 = tr.o4 = 0;
                        tr.w1 = tr.w2 = tr.w3 = tr.w4 = 0.f;
                    }
                    tapbuf[tp & 1][row] = tr;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kTcProducerThreads) : "memory");
            }
            const float *xc = xt + c0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int4 o = *reinterpret_cast<const int4 *>(&tapbuf[tp & 1][r0 + 32 * i]);
                raw[4 * i + 0] = __ldg(reinterpret_cast<const float4 *>(xc + o.x));
                raw[4 * i + 1] = __ldg(reinterpret_cast<const float4 *>(xc + o.y));
                raw[4 * i + 2] = __ldg(reinterpret_cast<const float4 *>(xc + o.z));
                raw[4 * i + 3] = __ldg(reinterpret_cast<const float4 *>(xc + o.w));
            }
        };
        prefetch(0);
        for (int kb = 0; kb < nkb; ++kb) {
            const int st = kb % stages, it = kb / stages;
            const int tp = kb / ncb;
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 w = *reinterpret_cast<const float4 *>(&tapbuf[tp & 1][r0 + 32 * i].w1);
                const float4 a1 = raw[4 * i], a2 = raw[4 * i + 1], a3 = raw[4 * i + 2], a4 = raw[4 * i + 3];
                v[i].x = fmaf(w.w, a4.x, fmaf(w.z, a3.x, fmaf(w.y, a2.x, w.x * a1.x)));
                v[i].y = fmaf(w.w, a4.y, fmaf(w.z, a3.y, fmaf(w.y, a2.y, w.x * a1.y)));
                v[i].z = fmaf(w.w, a4.z, fmaf(w.z, a3.z, fmaf(w.y, a2.z, w.x * a1.z)));
                v[i].w = fmaf(w.w, a4.w, fmaf(w.z, a3.w, fmaf(w.y, a2.w, w.x * a1.w)));
            }
            if (kb + 1 < nkb) prefetch(kb + 1);          // next block's loads fly while this one is stored
            if (it > 0) mbar_wait(&empty_bar[st], (uint32_t)((it - 1) & 1));
            unsigned char *sa = tiles + (size_t)st * stage_bytes;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t off = sw128(r0 + 32 * i, c4);
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
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[st]);   // one arrival per producer warp
        }

        // ================= epilogue: TMEM -> registers -> NCHW =================
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const int lg = warp & 3, chalf = warp >> 2;          // TMEM lane group of this warp, column half
        const bool affine = s.flags & SIDE_DCN_FUSE_AFFINE, relu = s.flags & SIDE_DCN_FUSE_RELU;
        float *yp = a.y + (size_t)b * s.Cout * s.P + p;
        const int cbeg = chalf * (N / 2), cend = cbeg + N / 2;
        for (int c = cbeg; c < cend; c += 8) {
            float acc[8];
            tc_ld8(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)c, acc);
            if (SPLIT) {                      // cross-term accumulator (see the MMA issuer)
                float accx[8];
                tc_ld8(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(N + c), accx);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += accx[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = c + j;
                float o = acc[j] + (a.bias ? __ldg(a.bias + n) : 0.f);
                if (affine) o = fmaf(o, __ldg(a.scale + n), __ldg(a.shift + n));
                if (relu) o = fmaxf(o, 0.f);
                if (pix_ok) yp[(size_t)n * s.P] = o;
            }
        }
    } else if (warp == 8) {
        // ================= MMA issuer =================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % stages, it = kb / stages;
                mbar_wait(&full_bar[st], (uint32_t)(it & 1));
                tc_fence_after();
                const uint32_t sa = smem_u32(tiles + (size_t)st * stage_bytes);
                const uint32_t sb = sa + a_bytes;
#pragma unroll
                for (int k = 0; k < kTcBK / 8; ++k) {
                    const uint64_t a_hi = tc_smem_desc(sa + k * 32), b_hi = tc_smem_desc(sb + k * 32);
                    tc_mma_tf32(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                    if (SPLIT) {
                        // The two small cross terms go to a SECOND accumulator (columns N..2N): the tensor core adds
                        // into fp32 with truncation, ~0.5 ulp of the accumulator per MMA, so keeping them out of the
                        // main accumulator cuts that systematic bias 3x; the epilogue adds the two.
                        const uint64_t a_lo = tc_smem_desc(sa + kATileBytes + k * 32);
                        const uint64_t b_lo = tc_smem_desc(sb + b_part_bytes + k * 32);
                        tc_mma_tf32(tmem_d + (uint32_t)N, a_lo, b_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        tc_mma_tf32(tmem_d + (uint32_t)N, a_hi, b_lo, idesc, 1u);
                    }
                }
                tc_commit(&empty_bar[st]);     // smem stage reusable once these MMAs have read it
            }
            tc_commit(&tmem_full_bar);         // accumulator complete
        }
        __syncwarp();
    } else {
        // ================= weight loader (TMA bulk copies) =================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % stages, it = kb / stages;
                if (it > 0) mbar_wait(&empty_bar[st], (uint32_t)((it - 1) & 1));
                mbar_expect_tx(&full_bar[st], b_bytes);
                bulk_g2s(tiles + (size_t)st * stage_bytes + a_bytes, wp + (size_t)kb * (b_bytes / 4), b_bytes, &full_bar[st]);
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
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

static size_t tc_weight_bytes(int Cin, int Cout, int KK, int flags)
{
    const bool split = (flags & SIDE_DCN_PREC_MASK) == SIDE_DCN_PREC_3XTF32;
    return (sizeof(float) * (split ? 2 : 1) * (size_t)Cin * Cout * KK + 255) & ~(size_t)255;
}

// workspace = weight tiles + the channels-last copy of x
size_t dcn_fwd_tc_ws_bytes(int B, int Cin, int H, int W, int Cout, int KK, int flags)
{
    return tc_weight_bytes(Cin, Cout, KK, flags) + sizeof(float) * (size_t)B * Cin * H * W;
}

int dcn_fwd_tc(const DcnFwdArgs &a, const float *w, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const DcnShape &s = a.s;
    const bool split = (s.flags & SIDE_DCN_PREC_MASK) == SIDE_DCN_PREC_3XTF32;
    if (!dcn_fwd_tc_supported(s.Cin, s.Cout, s.dg)) {
        set_error("side_dcn_fwd: tcgen05 path needs dg == 1, Cin %% 32 == 0, Cout %% 16 == 0, 16 <= Cout <= 256");
        return SIDE_ERR_UNSUPPORTED;
    }
    if (ws_bytes < dcn_fwd_tc_ws_bytes(s.B, s.Cin, s.H, s.W, s.Cout, s.KK, s.flags)) {
        set_error("side_dcn_fwd: workspace too small for the tcgen05 weight tiles + NHWC copy");
        return SIDE_ERR_WORKSPACE;
    }
    if ((long long)s.B * s.H * s.W * s.Cin >= (1ll << 31)) {
        set_error("side_dcn_fwd: input too large for 32-bit gather offsets");
        return SIDE_ERR_UNSUPPORTED;
    }
    float *wp = reinterpret_cast<float *>(ws);
    float *xt = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(ws) + tc_weight_bytes(s.Cin, s.Cout, s.KK, s.flags));
    int rc = launch_nchw_to_nhwc(a.x, xt, s.B, s.Cin, s.H * s.W, st);
    if (rc) return rc;
    if ((rc = launch_tc_weight_prep(w, wp, s.Cout, s.Cin, s.KK, split ? 1 : 0, st))) return rc;

    const uint32_t stage_bytes = (split ? 2u : 1u) * (kATileBytes + (uint32_t)s.Cout * 128u);
    int stages = (int)((200u * 1024u) / stage_bytes);
    stages = std::max(2, std::min(stages, kTcMaxStages));
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < (split ? 2 : 1) * s.Cout) tmem_cols <<= 1;
    // instruction descriptor: D = fp32, A = B = tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = tc_idesc_tf32(kTcBM, s.Cout);
    const unsigned grid = (unsigned)ceil_div((long long)s.B * s.P, kTcBM);
    if (split) {
        if ((rc = set_smem_attr((const void *)dcn_fwd_tc_kernel<true>, smem))) return rc;
        dcn_fwd_tc_kernel<true><<<grid, kTcThreads, smem, st>>>(a, wp, xt, stages, idesc, tmem_cols);
    } else {
        if ((rc = set_smem_attr((const void *)dcn_fwd_tc_kernel<false>, smem))) return rc;
        dcn_fwd_tc_kernel<false><<<grid, kTcThreads, smem, st>>>(a, wp, xt, stages, idesc, tmem_cols);
    }
    SIDE_LAUNCH_CHECK("dcn_fwd_tc_kernel");
    return SIDE_OK;
}

}  // namespace side
