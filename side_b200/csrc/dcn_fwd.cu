// dcn_fwd.cu -- DCNv2 modulated deformable convolution forward, fp32 SIMT path.
//
// Reference: dcn_v2_cuda_forward (DCNv2/src/cuda/dcn_v2_cuda.cu:43-173) = rank-1 bias GEMM + im2col kernel that
// MATERIALISES columns[B, Cin*kh*kw, Ho*Wo] in HBM (8.8 - 70.8 MB per image per layer) + batched SGEMM, with six
// cudaMalloc/cudaFree and a pointer-array kernel per call.
//
// Here: one fused implicit-GEMM kernel.  M = B*Ho*Wo output pixels (batch folded into M so small layers still fill
// the machine), N = Cout, K = kh*kw*Cin in TAP-MAJOR order (k = tap*Cin + c): the bilinear sample geometry
// (4 offsets + 4 weights + mask) of a pixel depends only on the tap, so it is computed once per tap and reused
// for every input channel.  Each thread owns one pixel of the 128-pixel tile for the gather (coalesced along W),
// writes its 16-channel slice of the A tile to shared memory; weights are re-laid-out once per call to
// [tap][Cin][Cout] so B tiles are read with coalesced 16-byte loads.  8x8 register micro-tiles, register-staged
// double buffering (the gather loads for block k+1 are in flight while block k is multiplied).
// Epilogue fuses bias, optional eval-mode BatchNorm affine and ReLU (DeformConv, feature_extraction_dla34.py:345-357).
#include <algorithm>
#include "dcn_common.cuh"

namespace side {

constexpr int kBM = 128;   // pixels per CTA
constexpr int kBN = 64;    // output channels per CTA
constexpr int kCK = 16;    // input channels per k-block
constexpr int kDcnThreads = 128;

// w [Cout, Cin, KK] -> wt [KK, Cin, Cout]
__global__ void dcn_weight_relayout_kernel(const float *__restrict__ w, float *__restrict__ wt, int Cout, int Cin, int KK)
{
    const long long n = (long long)Cout * Cin * KK;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(i % Cout);
        const long long r = i / Cout;
        const int c = (int)(r % Cin), t = (int)(r / Cin);
        wt[i] = __ldg(w + ((size_t)o * Cin + c) * KK + t);
    }
}

__global__ void __launch_bounds__(kDcnThreads, 3) dcn_fwd_simt_kernel(DcnFwdArgs a)
{
    __shared__ __align__(16) float As[2][kCK][kBM];
    __shared__ __align__(16) float Bs[2][kCK][kBN];
    const DcnShape &s = a.s;
    const int t = threadIdx.x;
    const int HWin = s.H * s.W;
    const long long Mtot = (long long)s.B * s.P;
    const long long gp = (long long)blockIdx.x * kBM + t;   // this thread's gather pixel
    const int n0 = blockIdx.y * kBN;
    const bool pix_ok = gp < Mtot;
    const int b = pix_ok ? (int)(gp / s.P) : 0;
    const int p = pix_ok ? (int)(gp - (long long)b * s.P) : 0;
    const int ho = p / s.Wo, wo = p - ho * s.Wo;
    const float *xb = a.x + (size_t)b * s.Cin * HWin;
    const int cpg = s.Cin / s.dg;
    const int ncb = (s.Cin + kCK - 1) / kCK;       // channel blocks per tap
    const int nkb = s.KK * ncb;

    const int tx = t & 15, ty = t >> 4;             // micro-tile: pixels tx*8..+8, couts ty*8..+8
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    DcnTap tap{};
    int cur_tap = -1, cur_g = -1;
    float av[kCK];
    float4 bv[2];

    auto load_block = [&](int kb) {
        const int tp = kb / ncb, c0 = (kb - tp * ncb) * kCK;
        const int g = c0 / cpg;
        if (tp != cur_tap || g != cur_g) {
            cur_tap = tp; cur_g = g;
            if (pix_ok) tap = dcn_tap(s, a.offset, a.mask, b, g, tp, ho, wo);
            else { tap.o1 = tap.o2 = tap.o3 = tap.o4 = 0; tap.w1 = tap.w2 = tap.w3 = tap.w4 = 0.f; tap.m = 0.f; }
        }
#pragma unroll
        for (int k = 0; k < kCK; ++k) {
            const int c = c0 + k;
            float v = 0.f;
            if (c < s.Cin) {
                const float *xc = xb + (size_t)c * HWin;
                v = tap.w1 * __ldg(xc + tap.o1) + tap.w2 * __ldg(xc + tap.o2) + tap.w3 * __ldg(xc + tap.o3) +
                    tap.w4 * __ldg(xc + tap.o4);
            }
            av[k] = v * tap.m;
        }
        // B tile: rows c0..c0+15 of wt[tp], columns n0..n0+63 -> 256 float4, 2 per thread
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int q = t + r * kDcnThreads;      // 0..255
            const int kk = q >> 4, nn = (q & 15) << 2;
            const int c = c0 + kk, n = n0 + nn;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < s.Cin) {
                const float *wp = a.wt + ((size_t)tp * s.Cin + c) * s.Cout + n;
                if (n + 3 < s.Cout && (s.Cout & 3) == 0) v = __ldg(reinterpret_cast<const float4 *>(wp));
                else {
                    if (n + 0 < s.Cout) v.x = __ldg(wp + 0);
                    if (n + 1 < s.Cout) v.y = __ldg(wp + 1);
                    if (n + 2 < s.Cout) v.z = __ldg(wp + 2);
                    if (n + 3 < s.Cout) v.w = __ldg(wp + 3);
                }
            }
            bv[r] = v;
        }
    };
    auto store_block = [&](int buf) {
#pragma unroll
        for (int k = 0; k < kCK; ++k) As[buf][k][t] = av[k];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int q = t + r * kDcnThreads;
            *reinterpret_cast<float4 *>(&Bs[buf][q >> 4][(q & 15) << 2]) = bv[r];
        }
    };

    load_block(0);
    store_block(0);
    __syncthreads();
    for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        if (kb + 1 < nkb) load_block(kb + 1);
#pragma unroll
        for (int k = 0; k < kCK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][tx * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][tx * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][ty * 8]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][ty * 8 + 4]);
            const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float br[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        if (kb + 1 < nkb) store_block(buf ^ 1);
        __syncthreads();
    }

    // epilogue: this thread holds pixels m0+tx*8..+8 x couts n0+ty*8..+8
    const long long m0 = (long long)blockIdx.x * kBM + tx * 8;
    const bool affine = s.flags & SIDE_DCN_FUSE_AFFINE, relu = s.flags & SIDE_DCN_FUSE_RELU;
    const int ob = (int)(m0 / s.P);
    const int op = (int)(m0 - (long long)ob * s.P);
    const bool vec = (m0 + 7 < Mtot) && (op + 7 < s.P) && ((s.P & 3) == 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = n0 + ty * 8 + j;
        if (n >= s.Cout) break;
        const float bj = a.bias ? __ldg(a.bias + n) : 0.f;
        const float sc = affine ? __ldg(a.scale + n) : 1.f, sf = affine ? __ldg(a.shift + n) : 0.f;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = acc[i][j] + bj;
            if (affine) v = fmaf(v, sc, sf);
            if (relu) v = fmaxf(v, 0.f);
            o[i] = v;
        }
        if (vec) {
            float *yp = a.y + ((size_t)ob * s.Cout + n) * s.P + op;
            *reinterpret_cast<float4 *>(yp) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4 *>(yp + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long m = m0 + i;
                if (m < Mtot) {
                    const int bb = (int)(m / s.P);
                    const int pp = (int)(m - (long long)bb * s.P);
                    a.y[((size_t)bb * s.Cout + n) * s.P + pp] = o[i];
                }
            }
        }
    }
}

// dcn_fwd_tc.cu (tcgen05 / TMEM path)
int dcn_fwd_tc(const DcnFwdArgs &a, const float *w, void *ws, size_t ws_bytes, cudaStream_t st, const float *x_nhwc = nullptr);
size_t dcn_fwd_tc_ws_bytes(int B, int Cin, int H, int W, int Cout, int KK, int flags);
bool dcn_fwd_tc_supported(int Cin, int Cout, int dg);

}  // namespace side

using namespace side;

extern "C" size_t side_dcn_fwd_ws_bytes(int B, int Cin, int H, int W, int Cout, int kh, int kw, int flags)
{
    if (Cin <= 0 || Cout <= 0 || kh <= 0 || kw <= 0) return 0;
    const size_t simt = sizeof(float) * (size_t)Cin * Cout * kh * kw;
    if ((flags & SIDE_DCN_PREC_MASK) != SIDE_DCN_PREC_FP32 && dcn_fwd_tc_supported(Cin, Cout, 1))
        return std::max(simt, dcn_fwd_tc_ws_bytes(B, Cin, H, W, Cout, kh * kw, flags));
    return simt;
}

extern "C" int side_dcn_fwd(const float *x, const float *offset, const float *mask, const float *w, const float *bias,
                            const float *scale, const float *shift, float *y, int B, int Cin, int H, int W, int Cout,
                            int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int dg, long long offset_bs,
                            long long mask_bs, int flags, void *ws, size_t ws_bytes, void *stream)
{
    DcnFwdArgs a{};
    int rc = dcn_fill_shape(a.s, B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dg, offset_bs, mask_bs, flags);
    if (rc) return rc;
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(offset); SIDE_REQUIRE_DEV(mask); SIDE_REQUIRE_DEV(w); SIDE_REQUIRE_DEV(y);
    if (bias) SIDE_REQUIRE_DEV(bias);
    if (flags & SIDE_DCN_FUSE_AFFINE) { SIDE_REQUIRE_DEV(scale); SIDE_REQUIRE_DEV(shift); }
    const size_t need = side_dcn_fwd_ws_bytes(B, Cin, H, W, Cout, kh, kw, flags);
    if (ws == nullptr || ws_bytes < need) {
        set_error("side_dcn_fwd: workspace too small (%zu < %zu bytes)", ws_bytes, need);
        return SIDE_ERR_WORKSPACE;
    }
    SIDE_REQUIRE_DEV(ws);
    a.x = x; a.offset = offset; a.mask = mask; a.bias = bias; a.scale = scale; a.shift = shift; a.y = y;
    cudaStream_t st = (cudaStream_t)stream;
    // tensor-core precisions fall back to the (more accurate) fp32 SIMT kernel for shapes tcgen05 cannot tile
    if ((flags & SIDE_DCN_PREC_MASK) != SIDE_DCN_PREC_FP32 && dcn_fwd_tc_supported(Cin, Cout, dg))
        return dcn_fwd_tc(a, w, ws, ws_bytes, st);

    SIDE_REQUIRE((Cin / dg) % kCK == 0 || dg == 1, "side_dcn_fwd: with deformable_groups>1, Cin/dg must be a multiple of %d",
                 kCK);
    float *wt = reinterpret_cast<float *>(ws);
    const long long nW = (long long)Cout * Cin * kh * kw;
    dcn_weight_relayout_kernel<<<(unsigned)min((long long)1184, (nW + 255) / 256), 256, 0, st>>>(w, wt, Cout, Cin, kh * kw);
    SIDE_LAUNCH_CHECK("dcn_weight_relayout_kernel");
    a.wt = wt;
    dim3 grid(ceil_div((long long)B * a.s.P, kBM), ceil_div(Cout, kBN));
    dcn_fwd_simt_kernel<<<grid, kDcnThreads, 0, st>>>(a);
    SIDE_LAUNCH_CHECK("dcn_fwd_simt_kernel");
    return SIDE_OK;
}

// Channels-last front end of the tcgen05 path: x is ALREADY [B, H, W, Cin] and the offset / mask-logit tensor is the
// channels-last output [B, Ho, Wo, om_ld] of the offset convolution (channels 0 .. 2KK-1 = interleaved (dy, dx),
// 2KK .. 3KK-1 = mask logits), e.g. produced by side_conv3d_tc_fwd.  Saves the NHWC staging copy and reads all 27
// values of a pixel from one 128-byte row.
extern "C" int side_dcn_fwd_cl(const float *x_nhwc, const float *om_cl, int om_ld, const float *w, const float *bias,
                               const float *scale, const float *shift, float *y, int B, int Cin, int H, int W, int Cout, int kh,
                               int kw, int sh, int sw, int ph, int pw, int dh, int dw, int flags, void *ws, size_t ws_bytes,
                               void *stream)
{
    DcnFwdArgs a{};
    flags |= SIDE_DCN_MASK_IS_LOGIT;
    int rc = dcn_fill_shape(a.s, B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, 1, 1, 1, flags);
    if (rc) return rc;
    SIDE_REQUIRE(om_ld >= 3 * kh * kw, "side_dcn_fwd_cl: om_ld=%d < 3*kh*kw", om_ld);
    SIDE_REQUIRE((flags & SIDE_DCN_PREC_MASK) != SIDE_DCN_PREC_FP32 && dcn_fwd_tc_supported(Cin, Cout, 1),
                 "side_dcn_fwd_cl: only the tcgen05 precisions (3xTF32 / TF32) with Cin %% 32 == 0, Cout %% 16 == 0 are built");
    SIDE_REQUIRE_DEV(x_nhwc); SIDE_REQUIRE_DEV(om_cl); SIDE_REQUIRE_DEV(w); SIDE_REQUIRE_DEV(y);
    if (bias) SIDE_REQUIRE_DEV(bias);
    if (flags & SIDE_DCN_FUSE_AFFINE) { SIDE_REQUIRE_DEV(scale); SIDE_REQUIRE_DEV(shift); }
    const size_t need = side_dcn_fwd_ws_bytes(B, Cin, H, W, Cout, kh, kw, flags);
    if (ws == nullptr || ws_bytes < need) {
        set_error("side_dcn_fwd_cl: workspace too small (%zu < %zu bytes)", ws_bytes, need);
        return SIDE_ERR_WORKSPACE;
    }
    SIDE_REQUIRE_DEV(ws);
    a.s.offset_bs = a.s.mask_bs = (long long)a.s.P * om_ld;
    a.s.om_cs = 1; a.s.om_ps = om_ld;
    a.x = x_nhwc; a.offset = om_cl; a.mask = om_cl + 2 * a.s.KK; a.bias = bias; a.scale = scale; a.shift = shift; a.y = y;
    return dcn_fwd_tc(a, w, ws, ws_bytes, (cudaStream_t)stream, x_nhwc);
}
