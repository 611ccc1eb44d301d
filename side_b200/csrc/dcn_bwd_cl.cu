// dcn_bwd_cl.cu -- channels-last DCNv2 backward, the path side_dcn_bwd takes for the
// DLA-34 neck shapes (deformable_groups == 1, Cin % 64 == 0, P % 4 == 0).
//
// Reference: dcn_v2_cuda_backward (DCNv2/src/cuda/dcn_v2_cuda.cu:207-336) -- per SAMPLE: Sgemm (columns of the output
// gradient), col2im_coord (dcn_v2_im2col_cuda.cu:256-327), col2im (:197-254), im2col (:125-195), Sgemm (weights),
// Sgemv (bias).  The scalar kernel in dcn_bwd.cu keeps that thread mapping (one thread walks all channels of one
// (pixel, tap) and issues one 4-byte atomic per channel and corner): 141 M scalar atomics and 2.0 ms for the 64-channel
// 96x320 layer of two images.
//
// Here the column buffer is PIXEL-major, gcol[b*P + p][tap*Cin + c], and the input / its gradient are channels-last:
//   1. gcol = gy^T Wp with Wp[o][tap*Cin + c] = w[o][c][tap]: the tcgen05 3xTF32 kernel of conv_tc.cu run as a plain GEMM
//      (gy re-laid channels-last and split hi/lo, weights permuted + swizzled by one small kernel) when Cout % 32 == 0 and
//      B*P % 128 == 0, else a tiled fp32 SGEMM (64x128x16 tiles)
//   2. scatter kernel: a warp owns 32 pixels x one tap at a time; the sampling geometry of the 32 items goes through a
//      per-warp shared table, then the lanes of the warp spread over CHANNEL QUADS of one (Cin >= 128) or two (Cin == 64)
//      items: every corner read is one 16-byte load, the column is rewritten in place with the forward value, and
//      grad_input is accumulated with ONE 16-byte vector reduction (REDG.E.ADD.F32x4) per corner and channel quad --
//      4x fewer atomic operations, 32 lanes of parallelism along the channels instead of a serial channel loop.
//      grad_offset / grad_mask: lane-parallel partial sums, one butterfly per item, coalesced stores per tap.
//   3. gWp = gy col: split-K over the pixels on tcgen05 (dcn_gw_tc.cu: the scatter kernel emits the columns as fp16 pairs and the
//      pixel-major buffer is read in place as an MN-major B operand), or the fp32 SIMT tile kernel below when the shapes do not
//      tile (P % 64 != 0); 16-byte reductions on the small weight gradient, un-permuted into gw by a copy kernel
//   4. gx: channels-last accumulator -> NCHW
// Accumulation order is not deterministic -- as in the reference (DCNv2/README.md:47-62).
#include <algorithm>
#include "dcn_common.cuh"
#include "tc_common.cuh"

namespace side {

// ------------------------------------------------------------------------------------------------
// fp32 SGEMM, C[m][n] (+)= sum_k A(m,k) B(k,n);  B(k,n) = B[k*ldb + n];  A_MCONTIG ? A[k*lda + m] : A[m*lda + k]
//   grid = (ceil(N/128), ceil(M/64), batch*splits); split s covers k in [s*kchunk, min(K, (s+1)*kchunk))
//   requires 16-byte aligned rows (lda, ldb multiples of 4, M % 4 == 0 when A_MCONTIG, K % 4 == 0 otherwise, N % 4 == 0)
// ------------------------------------------------------------------------------------------------
constexpr int kSgM = 64, kSgN = 128, kSgK = 16;

template <bool A_MCONTIG, bool ATOMIC>
__global__ void __launch_bounds__(256) sgemm_tile_kernel(const float *__restrict__ A, const float *__restrict__ Bm,
                                                        float *__restrict__ C, int M, int N, int K, long long lda,
                                                        long long ldb, long long ldc, long long bsA, long long bsB,
                                                        long long bsC, int splits, int kchunk)
{
    __shared__ __align__(16) float As[kSgK][kSgM + 4];
    __shared__ __align__(16) float Bs[kSgK][kSgN + 4];
    const int t = threadIdx.x;
    const int batch = blockIdx.z / splits, split = blockIdx.z - batch * splits;
    A += (size_t)batch * bsA;
    Bm += (size_t)batch * bsB;
    C += (size_t)batch * bsC;
    const int m0 = blockIdx.y * kSgM, n0 = blockIdx.x * kSgN;
    const int kbeg = split * kchunk, kend = min(K, kbeg + kchunk);
    const int tx = t & 15, ty = t >> 4;

    // global -> register staging of one k-tile
    float4 ra, rb[2];
    auto load_tile = [&](int k0) {
        ra = make_float4(0.f, 0.f, 0.f, 0.f);
        if (A_MCONTIG) {
            const int k = t >> 4, mq = (t & 15) * 4;
            if (k0 + k < kend && m0 + mq < M) ra = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(k0 + k) * lda + m0 + mq));
        } else {
            const int m = t >> 2, kq = (t & 3) * 4;
            if (m0 + m < M && k0 + kq < kend) ra = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(m0 + m) * lda + k0 + kq));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int idx = t + r * 256, k = idx >> 5, nq = (idx & 31) * 4;
            rb[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + k < kend && n0 + nq < N) rb[r] = __ldg(reinterpret_cast<const float4 *>(Bm + (size_t)(k0 + k) * ldb + n0 + nq));
        }
    };
    auto store_tile = [&]() {
        if (A_MCONTIG) {
            const int k = t >> 4, mq = (t & 15) * 4;
            *reinterpret_cast<float4 *>(&As[k][mq]) = ra;
        } else {
            const int m = t >> 2, kq = (t & 3) * 4;
            As[kq][m] = ra.x; As[kq + 1][m] = ra.y; As[kq + 2][m] = ra.z; As[kq + 3][m] = ra.w;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int idx = t + r * 256, k = idx >> 5, nq = (idx & 31) * 4;
            *reinterpret_cast<float4 *>(&Bs[k][nq]) = rb[r];
        }
    };

    float acc[4][8] = {};
    if (kbeg < kend) load_tile(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += kSgK) {
        store_tile();
        __syncthreads();
        if (k0 + kSgK < kend) load_tile(k0 + kSgK);
#pragma unroll
        for (int k = 0; k < kSgK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[k][64 + tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = n0 + h * 64 + tx * 4;
            if (n >= N) continue;
            float *cp = C + (size_t)m * ldc + n;
            const float4 v = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
            if (ATOMIC) atomicAdd(reinterpret_cast<float4 *>(cp), v);
            else *reinterpret_cast<float4 *>(cp) = v;
        }
    }
}

// Wp[o][tap*Cin + c] = w[o][c][tap]   (TO_PERM)      /      w[o][c][tap] = Wp[o][tap*Cin + c]   (!TO_PERM)
template <bool TO_PERM>
__global__ void __launch_bounds__(256) dcn_weight_perm_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                                             int Cout, int Cin, int KK)
{
    const long long total = (long long)Cout * Cin * KK;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // index in the [o][tap][c] order
    if (i >= total) return;
    const int c = (int)(i % Cin);
    const long long r = i / Cin;
    const int tap = (int)(r % KK);
    const long long o = r / KK;
    const long long j = (o * Cin + c) * KK + tap;                           // index in the [o][c][tap] order
    if (TO_PERM) dst[i] = __ldg(src + j);
    else dst[j] = __ldg(src + i);
}

// Weights of the column GEMM for conv_tc_rows_gemm: n = tap*Cin + c (output column), k = o; n-tiles of Nt columns, each
// [k-block][hi|lo][Nt x 32] in the 128-byte-swizzled shared-memory image (see dcn_tc_weight_prep_kernel).
__global__ void __launch_bounds__(256) dcn_bwd_wprep_tc_kernel(const float *__restrict__ w, float *__restrict__ wp, int Cout,
                                                              int Cin, int KK, int Nt)
{
    const long long total = (long long)Cout * Cin * KK;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over (n, k), k fastest
    if (i >= total) return;
    const int k = (int)(i % Cout);
    const int n = (int)(i / Cout);
    const int tap = n / Cin, c = n - tap * Cin;
    const float v = __ldg(w + ((size_t)k * Cin + c) * KK + tap);
    const int nt = n / Nt, nl = n - nt * Nt, kb = k >> 5, kk = k & 31;
    const size_t tile = (size_t)nt * 2 * Nt * Cout + (size_t)kb * 2 * Nt * 32;
    const uint32_t off = (sw128(nl, kk >> 2) >> 2) + (kk & 3);
    const float hi = tf32_hi(v);
    wp[tile + off] = hi;
    wp[tile + (size_t)Nt * 32 + off] = v - hi;
}

int conv_tc_rows_gemm(const float *x_hi, const float *x_lo, const float *wp, float *y, long long ldy, long long rows, int K,
                      int Ncols, int Nt, cudaStream_t st);
// weight gradient on tcgen05 (dcn_gw_tc.cu)
bool dcn_gw_tc_supported(int Cout, int Cin, int KK, int P, long long rows, bool partial_ok = false);
int dcn_gw_tc_split_gy(const float *gy, void *gy_pairs, int B, int Cout, int P, cudaStream_t st);
int dcn_gw_tc_run(const void *gy_pairs, const void *col_hi, const void *col_lo, float *gw, int B, int b0, int nb, int Cout, int Kp,
                  int P, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// scatter kernel
// ------------------------------------------------------------------------------------------------
struct DcnBwdClArgs {
    const float *x_cl;             // [B][H][W][Cin]
    const float *offset, *mask;    // as in the forward (batch strides / layout in s)
    float *gcol;                   // [nb*P][KK*Cin]  in: W^T gy, out: forward columns
    void *col_hi, *col_lo;         // non-null: the forward columns go here as fp16 pairs [nb*P][KK*Cin] (tcgen05 weight GEMM)
    float *gx_cl;                  // [B][H][W][Cin] accumulator (zeroed) or null
    float *goffset, *gmask;        // NCHW planes (batch strides s.offset_bs / s.mask_bs) or null
    DcnShape s;
    int b0, nb;                    // sample chunk
    int tiles_w, tiles_h, tg;      // 8x16-pixel tiles, taps per CTA (1, 3 or 9 for 3x3)
};

constexpr int kScTileH = 8, kScTileW = 16;

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// LPI = lanes per item (channel quads handled at once): 16 for Cin == 64 (two items per step), 32 otherwise
template <int LPI>
__global__ void __launch_bounds__(128) dcn_bwd_scatter_cl_kernel(DcnBwdClArgs a)
{
    __shared__ uint4 tab[4][32][3];       // per warp and item: {o1..o4 (x Cin)}, {w1..w4}, {lh, lw, m, valid bits}
    __shared__ float red[4][32][3];       // per warp and item: grad_mask, grad_offset_h, grad_offset_w
    const DcnShape &s = a.s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntg = s.KK / a.tg;
    int r = blockIdx.x;
    const int tgi = r % ntg; r /= ntg;
    const int twi = r % a.tiles_w; r /= a.tiles_w;
    const int thi = r % a.tiles_h;
    const int bl = r / a.tiles_h, b = a.b0 + bl;
    const int th0 = thi * kScTileH + 2 * warp, tw0 = twi * kScTileW;
    const int ho = th0 + (lane >> 4), wo = tw0 + (lane & 15);
    const bool pix_ok = ho < s.Ho && wo < s.Wo;
    const int p = ho * s.Wo + wo;
    const int Cin = s.Cin, Kp = s.KK * Cin;
    constexpr int SIM = 32 / LPI;
    const int sub = lane / LPI, cl = lane % LPI;
    const int CI = (Cin + 4 * LPI - 1) / (4 * LPI);
    const float *xb = a.x_cl + (size_t)b * s.H * s.W * Cin;
    float *gxb = a.gx_cl ? a.gx_cl + (size_t)b * s.H * s.W * Cin : nullptr;
    float *gcb = a.gcol + (size_t)bl * s.P * Kp;

    for (int t = 0; t < a.tg; ++t) {
        const int tap = tgi * a.tg + t;
        uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0, q2 = q0;
        float m_own = 0.f;
        if (pix_ok) {
            DcnTapGrad tg;
            const DcnTap tp = dcn_tap(s, a.offset, a.mask, b, 0, tap, ho, wo, &tg);
            q0 = make_uint4((unsigned)(tp.o1 * Cin), (unsigned)(tp.o2 * Cin), (unsigned)(tp.o3 * Cin), (unsigned)(tp.o4 * Cin));
            q1 = make_uint4(__float_as_uint(tp.w1), __float_as_uint(tp.w2), __float_as_uint(tp.w3), __float_as_uint(tp.w4));
            const unsigned vb = (tg.v1 ? 1u : 0u) | (tg.v2 ? 2u : 0u) | (tg.v3 ? 4u : 0u) | (tg.v4 ? 8u : 0u) | 16u;
            q2 = make_uint4(__float_as_uint(tg.lh), __float_as_uint(tg.lw), __float_as_uint(tp.m), vb);
            m_own = tp.m;
        }
        __syncwarp();
        tab[warp][lane][0] = q0; tab[warp][lane][1] = q1; tab[warp][lane][2] = q2;
        __syncwarp();

#pragma unroll 2
        for (int j = 0; j < 32; j += SIM) {
            const int it = j + sub;
            const uint4 o = tab[warp][it][0], wq = tab[warp][it][1], gq = tab[warp][it][2];
            float mv = 0.f, vh = 0.f, vw = 0.f;
            if (gq.w & 16u) {
                const float w1 = __uint_as_float(wq.x), w2 = __uint_as_float(wq.y), w3 = __uint_as_float(wq.z),
                            w4 = __uint_as_float(wq.w);
                const float lh = __uint_as_float(gq.x), lw = __uint_as_float(gq.y), m = __uint_as_float(gq.z);
                const float hh = 1.f - lh, hw = 1.f - lw;
                // d val / d h = a1 v1 + a2 v2 + a3 v3 + a4 v4,   d val / d w = b1 v1 + ... (dcn_v2_im2col_cuda.cu:89-122)
                const float a1 = (gq.w & 1u) ? -hw : 0.f, a2 = (gq.w & 2u) ? -lw : 0.f, a3 = (gq.w & 4u) ? hw : 0.f,
                            a4 = (gq.w & 8u) ? lw : 0.f;
                const float b1 = (gq.w & 1u) ? -hh : 0.f, b2 = (gq.w & 2u) ? hh : 0.f, b3 = (gq.w & 4u) ? -lh : 0.f,
                            b4 = (gq.w & 8u) ? lh : 0.f;
                const int pi = (th0 + (it >> 4)) * s.Wo + tw0 + (it & 15);
                float *gc = gcb + (size_t)pi * Kp + (size_t)tap * Cin;
                for (int ci = 0; ci < CI; ++ci) {
                    const int c0 = (ci * LPI + cl) * 4;
                    if (c0 >= Cin) break;                                   // Cin = 64 * odd: last pass is half populated
                    const float4 g = *reinterpret_cast<const float4 *>(gc + c0);
                    const float4 v1 = __ldg(reinterpret_cast<const float4 *>(xb + o.x + c0));
                    const float4 v2 = __ldg(reinterpret_cast<const float4 *>(xb + o.y + c0));
                    const float4 v3 = __ldg(reinterpret_cast<const float4 *>(xb + o.z + c0));
                    const float4 v4 = __ldg(reinterpret_cast<const float4 *>(xb + o.w + c0));
                    float4 val, top;
                    val.x = w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
                    val.y = w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
                    val.z = w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
                    val.w = w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
                    if (a.col_hi) {
                        uint32_t h0, l0, h1, l1;
                        f16_split2(val.x * m, val.y * m, h0, l0);
                        f16_split2(val.z * m, val.w * m, h1, l1);
                        const size_t e4 = (((size_t)bl * s.P + pi) * Kp + (size_t)tap * Cin + c0) >> 2;
                        reinterpret_cast<uint2 *>(a.col_hi)[e4] = make_uint2(h0, h1);
                        reinterpret_cast<uint2 *>(a.col_lo)[e4] = make_uint2(l0, l1);
                    } else {
                        *reinterpret_cast<float4 *>(gc + c0) = make_float4(val.x * m, val.y * m, val.z * m, val.w * m);
                    }
                    mv += g.x * val.x + g.y * val.y + g.z * val.z + g.w * val.w;
                    top.x = g.x * m; top.y = g.y * m; top.z = g.z * m; top.w = g.w * m;
                    vh += (a1 * v1.x + a2 * v2.x + a3 * v3.x + a4 * v4.x) * top.x + (a1 * v1.y + a2 * v2.y + a3 * v3.y + a4 * v4.y) * top.y +
                          (a1 * v1.z + a2 * v2.z + a3 * v3.z + a4 * v4.z) * top.z + (a1 * v1.w + a2 * v2.w + a3 * v3.w + a4 * v4.w) * top.w;
                    vw += (b1 * v1.x + b2 * v2.x + b3 * v3.x + b4 * v4.x) * top.x + (b1 * v1.y + b2 * v2.y + b3 * v3.y + b4 * v4.y) * top.y +
                          (b1 * v1.z + b2 * v2.z + b3 * v3.z + b4 * v4.z) * top.z + (b1 * v1.w + b2 * v2.w + b3 * v3.w + b4 * v4.w) * top.w;
                    if (gxb) {
                        if (w1 != 0.f) red_add_v4(gxb + o.x + c0, w1 * top.x, w1 * top.y, w1 * top.z, w1 * top.w);
                        if (w2 != 0.f) red_add_v4(gxb + o.y + c0, w2 * top.x, w2 * top.y, w2 * top.z, w2 * top.w);
                        if (w3 != 0.f) red_add_v4(gxb + o.z + c0, w3 * top.x, w3 * top.y, w3 * top.z, w3 * top.w);
                        if (w4 != 0.f) red_add_v4(gxb + o.w + c0, w4 * top.x, w4 * top.y, w4 * top.z, w4 * top.w);
                    }
                }
            }
#pragma unroll
            for (int d = LPI / 2; d >= 1; d >>= 1) {
                mv += __shfl_xor_sync(0xffffffffu, mv, d);
                vh += __shfl_xor_sync(0xffffffffu, vh, d);
                vw += __shfl_xor_sync(0xffffffffu, vw, d);
            }
            if (cl == 0) { red[warp][it][0] = mv; red[warp][it][1] = vh; red[warp][it][2] = vw; }
        }
        __syncwarp();
        if (pix_ok) {
            if (a.goffset) {
                float *go = a.goffset + (size_t)b * s.offset_bs + (size_t)(2 * tap) * s.P + p;
                go[0] = red[warp][lane][1];
                go[s.P] = red[warp][lane][2];
            }
            if (a.gmask) {
                float mval = red[warp][lane][0];
                if (s.flags & SIDE_DCN_MASK_IS_LOGIT) mval *= m_own * (1.f - m_own);
                a.gmask[(size_t)b * s.mask_bs + (size_t)tap * s.P + p] = mval;
            }
        }
    }
}

// gbias[o] += sum over a slice of (b, p) of gy[b, o, p]; grid = (Cout, slices), gbias zeroed by the caller
__global__ void __launch_bounds__(256) dcn_bias_grad_split_kernel(const float *__restrict__ gy, float *__restrict__ gb, int B,
                                                                 int Cout, int P, int slices)
{
    __shared__ float red[8];
    const int o = blockIdx.x, sl = blockIdx.y;
    const long long total4 = (long long)B * (P / 4);
    const long long per = (total4 + slices - 1) / slices;
    const long long beg = sl * per, end = min(total4, beg + per);
    float sum = 0.f;
    const int P4 = P / 4;
    for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
        const int b = (int)(i / P4), q = (int)(i - (long long)b * P4);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(gy + ((size_t)b * Cout + o) * P) + q);
        sum += (v.x + v.y) + (v.z + v.w);
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(gb + o, v);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool dcn_bwd_cl_supported(const DcnShape &s)
{
    return s.dg == 1 && s.Cin % 64 == 0 && s.P % 4 == 0 && s.Cout % 4 == 0 && s.om_ps == 1 &&
           (long long)s.H * s.W * s.Cin < (1ll << 31) && s.B <= 65535;
}

// floats: [Wp][gWp][x_cl][gx_cl] fixed, then the column buffer (per sample P*KK*Cin)
size_t dcn_bwd_cl_fixed_floats(const DcnShape &s)
{
    return 2 * (size_t)s.Cout * s.Cin * s.KK + 2 * (size_t)s.B * s.H * s.W * s.Cin;
}
// extra floats of the tensor-core column GEMM: swizzled hi/lo weights + channels-last hi/lo copy of gy
size_t dcn_bwd_cl_tc_floats(const DcnShape &s)
{
    return 2 * (size_t)s.Cout * s.Cin * s.KK + 2 * (size_t)s.B * s.P * s.Cout;
}
// extra floats of the tensor-core weight GEMM besides a second column buffer per sample: fp16 pairs of gy
// (+ 64 floats: the power-of-two range scale of gy, see dcn_gw_tc.cu)
size_t dcn_bwd_cl_gw_floats(const DcnShape &s) { return (size_t)s.B * s.P * s.Cout + 64; }

int dcn_bwd_cl_run(const DcnShape &s, const float *x, const float *offset, const float *mask, const float *w, const float *gy,
                   float *gx, float *goffset, float *gmask, float *gw, float *gbias, float *ws, size_t ws_floats,
                   cudaStream_t st)
{
    const int Cin = s.Cin, Cout = s.Cout, KK = s.KK, P = s.P, B = s.B, Kp = KK * Cin;
    const size_t wsz = (size_t)Cout * Kp, xsz = (size_t)B * s.H * s.W * Cin, per_sample = (size_t)P * Kp;
    float *Wp = ws, *gWp = Wp + wsz, *x_cl = gWp + wsz, *gx_cl = x_cl + xsz, *gcol = gx_cl + xsz;
    const size_t fixed = 2 * wsz + 2 * xsz;
    SIDE_REQUIRE(ws_floats >= fixed + per_sample, "side_dcn_bwd: workspace too small for the channels-last path");
    int rc;
    // tensor-core column GEMM when the shapes tile and the workspace has room for its operands next to >= 1 sample of columns
    const size_t tcf = dcn_bwd_cl_tc_floats(s);
    // n-tile of the column GEMM: 128 columns when Kp tiles by 128, else 64 (Cin = 64 * odd: Kp = 9 * Cin is a multiple of 64 only);
    // every 1536-column slice and the tail Kp % 1536 are then whole n-tiles and wtc (2 * wsz floats) holds exactly Kp / Nt of them
    const int Nt = (Cin >= 128 && Kp % 128 == 0) ? 128 : 64;
    const bool use_tc = !(s.flags & SIDE_DCN_BWD_SIMT_GEMM) && Cout % 32 == 0 && Kp % Nt == 0 && ws_floats >= fixed + tcf + per_sample;
    float *wtc = gcol, *gy_hi = wtc + 2 * wsz, *gy_lo = gy_hi + (size_t)B * P * Cout;
    size_t avail = ws_floats - fixed;
    if (use_tc) { gcol = gy_lo + (size_t)B * P * Cout; avail -= tcf; }
    int chunk = (int)std::min<size_t>((size_t)B, avail / per_sample);
    if (use_tc && ((size_t)chunk * P) % 128 != 0) {
        // chunks must cover whole 128-row tiles: shrink to the largest chunk that does, or give up the tensor-core GEMM
        while (chunk > 1 && (((size_t)chunk * P) % 128 != 0 || B % chunk != 0)) --chunk;
    }
    const bool tc = use_tc && ((size_t)chunk * P) % 128 == 0 && B % chunk == 0;
    if (use_tc && !tc) { gcol = wtc; avail = ws_floats - fixed; chunk = (int)std::min<size_t>((size_t)B, avail / per_sample); }
    // tensor-core weight GEMM: needs the columns as fp16 pairs (a second buffer of the same size per sample) and the gy pairs
    bool gwtc = false;
    float *gy_pairs = nullptr, *col_pairs = nullptr;
    if (gw && tc && dcn_gw_tc_supported(Cout, Cin, KK, P, (long long)chunk * P)) {
        const size_t gyf = dcn_bwd_cl_gw_floats(s);
        if (avail >= gyf + 2 * per_sample) {
            int c2 = (int)std::min<size_t>((size_t)B, (avail - gyf) / (2 * per_sample));
            while (c2 > 1 && (((size_t)c2 * P) % 128 != 0 || B % c2 != 0)) --c2;
            if (((size_t)c2 * P) % 128 == 0 && B % c2 == 0) {
                gwtc = true;
                chunk = c2;
                gy_pairs = gcol;                                     // [gy pairs][gcol chunk][column pairs chunk]
                gcol = gy_pairs + gyf;
                col_pairs = gcol + (size_t)chunk * per_sample;
            }
        }
    }
    if (gwtc && (rc = dcn_gw_tc_split_gy(gy, gy_pairs, B, Cout, P, st))) return rc;
    if (tc) {
        dcn_bwd_wprep_tc_kernel<<<ceil_div((long long)wsz, 256), 256, 0, st>>>(w, wtc, Cout, Cin, KK, Nt);
        SIDE_LAUNCH_CHECK("dcn_bwd_wprep_tc_kernel");
        if ((rc = side_ncdhw_to_cl_split(gy, nullptr, nullptr, gy_hi, gy_lo, B, Cout, P, 1, st))) return rc;
    }
    if ((rc = launch_nchw_to_nhwc(x, x_cl, B, Cin, s.H * s.W, st))) return rc;
    if (gx) SIDE_CUDA(cudaMemsetAsync(gx_cl, 0, sizeof(float) * xsz, st));
    if (gw) SIDE_CUDA(cudaMemsetAsync(gWp, 0, sizeof(float) * wsz, st));

    if (!tc) {
        dcn_weight_perm_kernel<true><<<ceil_div((long long)wsz, 256), 256, 0, st>>>(w, Wp, Cout, Cin, KK);
        SIDE_LAUNCH_CHECK("dcn_weight_perm_kernel");
    }
    DcnBwdClArgs a{};
    a.x_cl = x_cl; a.offset = offset; a.mask = mask; a.gcol = gcol; a.gx_cl = gx ? gx_cl : nullptr;
    a.goffset = goffset; a.gmask = gmask; a.s = s;
    a.tiles_w = ceil_div(s.Wo, kScTileW); a.tiles_h = ceil_div(s.Ho, kScTileH);
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        a.b0 = b0; a.nb = nb;
        const float *gyb = gy + (size_t)b0 * Cout * P;
        // 1. gcol[bl][p][k'] = sum_o gy[b][o][p] Wp[o][k']
        if (tc) {
            for (int col0 = 0; col0 < Kp; col0 += 1536) {
                const int ncols = std::min(1536, Kp - col0);
                if ((rc = conv_tc_rows_gemm(gy_hi + (size_t)b0 * P * Cout, gy_lo + (size_t)b0 * P * Cout, wtc + (size_t)col0 * 2 * Cout,
                                            gcol + col0, Kp, (long long)nb * P, Cout, ncols, Nt, st)))
                    return rc;
            }
        } else {
            dim3 grid(ceil_div(Kp, kSgN), ceil_div(P, kSgM), nb);
            sgemm_tile_kernel<true, false><<<grid, 256, 0, st>>>(gyb, Wp, gcol, P, Kp, Cout, P, Kp, Kp, (long long)Cout * P, 0,
                                                               (long long)P * Kp, 1, Cout);
            SIDE_LAUNCH_CHECK("sgemm_tile_kernel(gcol)");
        }
        // 2. grads w.r.t. input / offset / mask; gcol := forward columns
        {
            const long long tiles = (long long)nb * a.tiles_h * a.tiles_w;
            a.tg = 1;
            if (KK == 9) a.tg = tiles >= 1184 ? 9 : tiles * 3 >= 1184 ? 3 : 1;
            const long long ctas = tiles * (KK / a.tg);
            SIDE_REQUIRE(ctas < (1ll << 31), "side_dcn_bwd: grid too large");
            a.gcol = gcol;
            a.col_hi = gwtc ? reinterpret_cast<void *>(col_pairs) : nullptr;
            a.col_lo = gwtc ? reinterpret_cast<void *>(reinterpret_cast<__half *>(col_pairs) + (size_t)nb * P * Kp) : nullptr;
            if (Cin == 64) dcn_bwd_scatter_cl_kernel<16><<<(unsigned)ctas, 128, 0, st>>>(a);
            else dcn_bwd_scatter_cl_kernel<32><<<(unsigned)ctas, 128, 0, st>>>(a);
            SIDE_LAUNCH_CHECK("dcn_bwd_scatter_cl_kernel");
        }
        // 3. gWp[o][k'] += sum_p gy[b][o][p] col[bl][p][k']
        if (gw && gwtc) {
            if ((rc = dcn_gw_tc_run(gy_pairs, a.col_hi, a.col_lo, gWp, B, b0, nb, Cout, Kp, P, st))) return rc;
        } else if (gw) {
            const int tiles = ceil_div(Kp, kSgN) * ceil_div(Cout, kSgM) * nb;
            int splits = std::max(1, std::min(ceil_div(P, 256), ceil_div(592, tiles)));
            int kchunk = ceil_div(ceil_div(P, splits), kSgK) * kSgK;
            splits = ceil_div(P, kchunk);
            dim3 grid(ceil_div(Kp, kSgN), ceil_div(Cout, kSgM), nb * splits);
            sgemm_tile_kernel<false, true><<<grid, 256, 0, st>>>(gyb, gcol, gWp, Cout, Kp, P, P, Kp, Kp, (long long)Cout * P,
                                                               (long long)P * Kp, 0, splits, kchunk);
            SIDE_LAUNCH_CHECK("sgemm_tile_kernel(gW)");
        }
    }
    if (gw) {
        dcn_weight_perm_kernel<false><<<ceil_div((long long)wsz, 256), 256, 0, st>>>(gWp, gw, Cout, Cin, KK);
        SIDE_LAUNCH_CHECK("dcn_weight_perm_kernel");
    }
    if (gx && (rc = launch_nhwc_to_nchw(gx_cl, gx, B, Cin, s.H * s.W, st))) return rc;
    if (gbias) {
        SIDE_CUDA(cudaMemsetAsync(gbias, 0, sizeof(float) * Cout, st));
        const int slices = std::max(1, std::min(64, ceil_div(592, Cout)));
        dcn_bias_grad_split_kernel<<<dim3(Cout, slices), 256, 0, st>>>(gy, gbias, B, Cout, P, slices);
        SIDE_LAUNCH_CHECK("dcn_bias_grad_split_kernel");
    }
    return SIDE_OK;
}

}  // namespace side
