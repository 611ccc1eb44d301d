// volume.cu -- full-image disparity-sweep cost volumes (north_star item 1).
//
// The reference has no full-image builder; its stereo-head concat
// torch.cat((imgfea_left, imgfea_right), 1) (stereo_network_old.py:348) is the D = 1 slice of the concat
// volume.  Semantics (PSMNet / GwcNet convention) are fixed in include/side_b200.h.
//
// B200 design -- pure HBM-write-bound kernels, every input byte read from HBM once per sweep:
//   * concat: a CTA owns (b, c, TY consecutive rows).  Those rows are contiguous in NCHW, so the left and the
//     right tile each arrive with ONE TMA bulk copy (cp.async.bulk, SASS UBLKCP) signalled on an mbarrier;
//     the CTA then sweeps d = 0..D-1 and streams the masked / shifted rows with 16-byte evict-first stores
//     (TY*W*4 contiguous bytes per (d, half)).
//   * gwc: a CTA owns (b, g, y): the C/G rows of L and R are bulk-copied to shared memory, each thread keeps
//     its x column of L in registers across the disparity sweep.
#include "common.cuh"

namespace side {

constexpr int kTY = 8;          // rows per concat CTA
constexpr int kVolBlock = 256;

__global__ void __launch_bounds__(kVolBlock) concat_volume_fwd_kernel(const float *__restrict__ L,
                                                                     const float *__restrict__ R,
                                                                     float *__restrict__ vol, int C, int H, int W,
                                                                     int D, int ytiles, int use_tma)
{
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bar;
    const int yt = blockIdx.x % ytiles, bc = blockIdx.x / ytiles;  // bc = b*C + c
    const int b = bc / C, c = bc - b * C;
    const int y0 = yt * kTY, ny = min(kTY, H - y0), n = ny * W;
    float *Ls = sm, *Rs = sm + kTY * W;
    const float *lsrc = L + ((size_t)bc * H + y0) * W, *rsrc = R + ((size_t)bc * H + y0) * W;

    if (use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
            mbar_expect_tx(&bar, 2u * (uint32_t)n * 4u);
            bulk_g2s(Ls, lsrc, (uint32_t)n * 4u, &bar);
            bulk_g2s(Rs, rsrc, (uint32_t)n * 4u, &bar);
        }
        __syncthreads();
        mbar_wait(&bar, 0);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            Ls[i] = __ldg(lsrc + i);
            Rs[i] = __ldg(rsrc + i);
        }
        __syncthreads();
    }

    const size_t HW = (size_t)H * W;
    float *outL = vol + (((size_t)b * 2 * C + c) * D) * HW + (size_t)y0 * W;
    float *outR = vol + (((size_t)b * 2 * C + C + c) * D) * HW + (size_t)y0 * W;
    if ((W & 3) == 0) {
        const int W4 = W >> 2, n4 = n >> 2;
        for (int d = 0; d < D; ++d) {
            float4 *oL = reinterpret_cast<float4 *>(outL + (size_t)d * HW);
            float4 *oR = reinterpret_cast<float4 *>(outR + (size_t)d * HW);
            for (int i = threadIdx.x; i < n4; i += blockDim.x) {
                const int row = i / W4, x = (i - row * W4) << 2;
                const float *lr = Ls + row * W, *rr = Rs + row * W;
                float4 a = *reinterpret_cast<const float4 *>(lr + x), r;
                a.x = x + 0 >= d ? a.x : 0.f;
                a.y = x + 1 >= d ? a.y : 0.f;
                a.z = x + 2 >= d ? a.z : 0.f;
                a.w = x + 3 >= d ? a.w : 0.f;
                r.x = x + 0 >= d ? rr[x + 0 - d] : 0.f;
                r.y = x + 1 >= d ? rr[x + 1 - d] : 0.f;
                r.z = x + 2 >= d ? rr[x + 2 - d] : 0.f;
                r.w = x + 3 >= d ? rr[x + 3 - d] : 0.f;
                st_cs(oL + i, a);
                st_cs(oR + i, r);
            }
        }
    } else {
        for (int d = 0; d < D; ++d)
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int row = i / W, x = i - row * W;
                st_cs(outL + (size_t)d * HW + i, x >= d ? Ls[i] : 0.f);
                st_cs(outR + (size_t)d * HW + i, x >= d ? Rs[row * W + x - d] : 0.f);
            }
    }
}

// gL[b,c,y,x] = sum_{d<=x} g[b,c,d,y,x];  gR[b,c,y,x] = sum_{d : x+d<W} g[b,C+c,d,y,x+d]
__global__ void __launch_bounds__(kVolBlock) concat_volume_bwd_kernel(const float *__restrict__ g, float *__restrict__ gL,
                                                                     float *__restrict__ gR, int C, int H, int W, int D)
{
    const size_t HW = (size_t)H * W;
    const long long total = (long long)HW;  // per (b,c) plane
    const int bc = blockIdx.y;
    const int b = bc / C, c = bc - b * C;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W);
    const float *gl = g + (((size_t)b * 2 * C + c) * D) * HW + i;
    const float *gr = g + (((size_t)b * 2 * C + C + c) * D) * HW + i;
    float sl = 0.f, sr = 0.f;
    const int dl = min(D - 1, x), dr = min(D - 1, W - 1 - x);
    for (int d = 0; d <= dl; ++d) sl += __ldcs(gl + (size_t)d * HW);
    for (int d = 0; d <= dr; ++d) sr += __ldcs(gr + (size_t)d * HW + d);
    gL[(size_t)bc * HW + i] = sl;
    gR[(size_t)bc * HW + i] = sr;
}

constexpr int kMaxCpg = 32;

template <int CPG>
__global__ void __launch_bounds__(kVolBlock) gwc_volume_fwd_kernel(const float *__restrict__ L, const float *__restrict__ R,
                                                                  float *__restrict__ vol, int C, int H, int W, int D,
                                                                  int G, int cpg_rt, int use_tma)
{
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bar;
    const int cpg = CPG > 0 ? CPG : cpg_rt;
    const int y = blockIdx.x % H, bg = blockIdx.x / H;
    const int b = bg / G, gi = bg - b * G;
    float *Ls = sm, *Rs = sm + cpg * W;
    const size_t HW = (size_t)H * W;
    const float *lsrc = L + ((size_t)b * C + gi * cpg) * HW + (size_t)y * W;
    const float *rsrc = R + ((size_t)b * C + gi * cpg) * HW + (size_t)y * W;
    if (use_tma) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
            mbar_expect_tx(&bar, 2u * (uint32_t)(cpg * W) * 4u);
            for (int cc = 0; cc < cpg; ++cc) {
                bulk_g2s(Ls + cc * W, lsrc + (size_t)cc * HW, (uint32_t)W * 4u, &bar);
                bulk_g2s(Rs + cc * W, rsrc + (size_t)cc * HW, (uint32_t)W * 4u, &bar);
            }
        }
        __syncthreads();
        mbar_wait(&bar, 0);
    } else {
        for (int i = threadIdx.x; i < cpg * W; i += blockDim.x) {
            const int cc = i / W, x = i - cc * W;
            Ls[i] = __ldg(lsrc + (size_t)cc * HW + x);
            Rs[i] = __ldg(rsrc + (size_t)cc * HW + x);
        }
        __syncthreads();
    }
    const float inv = 1.0f / (float)cpg;
    float *out = vol + (((size_t)b * G + gi) * D) * HW + (size_t)y * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        float lv[CPG > 0 ? CPG : 1];
        if (CPG > 0) {
#pragma unroll
            for (int cc = 0; cc < CPG; ++cc) lv[cc] = Ls[cc * W + x];
        }
        for (int d = 0; d < D; ++d) {
            float s = 0.f;
            if (x >= d) {
                if (CPG > 0) {
#pragma unroll
                    for (int cc = 0; cc < CPG; ++cc) s = fmaf(lv[cc], Rs[cc * W + x - d], s);
                } else {
                    for (int cc = 0; cc < cpg; ++cc) s = fmaf(Ls[cc * W + x], Rs[cc * W + x - d], s);
                }
            }
            st_cs(out + (size_t)d * HW + x, s * inv);
        }
    }
}

// Register-tiled variant (W % 4 == 0, D % 4 == 0, compile-time channels per group): one thread produces a 4 (x) by 4 (d)
// block.  The 16 products of a channel need only the 7 right-view values R[x0-d-3 .. x0+3-d], fetched as two aligned
// 16-byte shared loads from a row that is zero-padded on the left by D + 4 columns (so x < d reads zeros and the mask
// needs no branch); outputs leave as 16-byte evict-first stores.  1 shared load per output instead of 8.
template <int CPG>
__global__ void __launch_bounds__(kVolBlock) gwc_volume_fwd_tiled_kernel(const float *__restrict__ L, const float *__restrict__ R,
                                                                        float *__restrict__ vol, int C, int H, int W, int D, int G)
{
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t bar;
    const int y = blockIdx.x % H, bg = blockIdx.x / H;
    const int b = bg / G, gi = bg - b * G;
    const int pad = D + 4, RW = W + pad;                 // right-view row stride (floats), pad % 4 == 0
    float *Ls = sm, *Rs = sm + CPG * W;
    const size_t HW = (size_t)H * W;
    const float *lsrc = L + ((size_t)b * C + gi * CPG) * HW + (size_t)y * W;
    const float *rsrc = R + ((size_t)b * C + gi * CPG) * HW + (size_t)y * W;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, 2u * (uint32_t)(CPG * W) * 4u);
        for (int cc = 0; cc < CPG; ++cc) {
            bulk_g2s(Ls + cc * W, lsrc + (size_t)cc * HW, (uint32_t)W * 4u, &bar);
            bulk_g2s(Rs + cc * RW + pad, rsrc + (size_t)cc * HW, (uint32_t)W * 4u, &bar);
        }
    }
    for (int i = threadIdx.x; i < CPG * pad; i += blockDim.x) Rs[(i / pad) * RW + (i % pad)] = 0.f;
    __syncthreads();
    mbar_wait(&bar, 0);
    const float inv = 1.0f / (float)CPG;
    float *out = vol + (((size_t)b * G + gi) * D) * HW + (size_t)y * W;
    const int nxq = W >> 2, ndq = D >> 2;
    for (int item = threadIdx.x; item < nxq * ndq; item += blockDim.x) {
        const int dq = item / nxq, xq = item - dq * nxq;
        const int x0 = 4 * xq, d0 = 4 * dq;
        float acc[4][4] = {};                             // [dd][xx]
#pragma unroll
        for (int cc = 0; cc < CPG; ++cc) {
            const float4 l = *reinterpret_cast<const float4 *>(Ls + cc * W + x0);
            const float *rp = Rs + cc * RW + pad + x0 - d0 - 4;      // 16-byte aligned: pad, x0, d0 are multiples of 4
            const float4 ra = *reinterpret_cast<const float4 *>(rp), rb = *reinterpret_cast<const float4 *>(rp + 4);
            const float r[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};   // r[k] = R[x0 - d0 - 4 + k]
            const float lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
            for (int dd = 0; dd < 4; ++dd)
#pragma unroll
                for (int xx = 0; xx < 4; ++xx) acc[dd][xx] = fmaf(lv[xx], r[4 + xx - dd], acc[dd][xx]);   // R[x0+xx-(d0+dd)]
        }
#pragma unroll
        for (int dd = 0; dd < 4; ++dd)
            st_cs(reinterpret_cast<float4 *>(out + (size_t)(d0 + dd) * HW + x0),
                  make_float4(acc[dd][0] * inv, acc[dd][1] * inv, acc[dd][2] * inv, acc[dd][3] * inv));
    }
}

// gL[c,y,x] = (1/cpg) sum_{d<=x} g[g(c),d,y,x] R[c,y,x-d];  gR[c,y,x] = (1/cpg) sum_{d: x+d<W} g[g(c),d,y,x+d] L[c,y,x+d]
__global__ void __launch_bounds__(kVolBlock) gwc_volume_bwd_kernel(const float *__restrict__ L, const float *__restrict__ R,
                                                                  const float *__restrict__ g, float *__restrict__ gL,
                                                                  float *__restrict__ gR, int C, int H, int W, int D,
                                                                  int G, int cpg)
{
    extern __shared__ __align__(16) float sm[];
    const int y = blockIdx.x % H, bg = blockIdx.x / H;
    const int b = bg / G, gi = bg - b * G;
    float *Ls = sm, *Rs = sm + cpg * W;
    const size_t HW = (size_t)H * W;
    const size_t fbase = ((size_t)b * C + gi * cpg) * HW + (size_t)y * W;
    for (int i = threadIdx.x; i < cpg * W; i += blockDim.x) {
        const int cc = i / W, x = i - cc * W;
        Ls[i] = __ldg(L + fbase + (size_t)cc * HW + x);
        Rs[i] = __ldg(R + fbase + (size_t)cc * HW + x);
    }
    __syncthreads();
    const float inv = 1.0f / (float)cpg;
    const float *gp = g + (((size_t)b * G + gi) * D) * HW + (size_t)y * W;
    for (int i = threadIdx.x; i < cpg * W; i += blockDim.x) {
        const int cc = i / W, x = i - cc * W;
        float sl = 0.f, sr = 0.f;
        const int dl = min(D - 1, x), dr = min(D - 1, W - 1 - x);
        for (int d = 0; d <= dl; ++d) sl = fmaf(__ldg(gp + (size_t)d * HW + x), Rs[cc * W + x - d], sl);
        for (int d = 0; d <= dr; ++d) sr = fmaf(__ldg(gp + (size_t)d * HW + x + d), Ls[cc * W + x + d], sr);
        gL[fbase + (size_t)cc * HW + x] = sl * inv;
        gR[fbase + (size_t)cc * HW + x] = sr * inv;
    }
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace side

using namespace side;

extern "C" int side_concat_volume_fwd(const float *L, const float *R, float *vol, int B, int C, int H, int W, int D,
                                      void *stream)
{
    SIDE_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "side_concat_volume_fwd: bad shape");
    SIDE_REQUIRE_DEV(L); SIDE_REQUIRE_DEV(R); SIDE_REQUIRE_DEV(vol);
    const int ytiles = ceil_div(H, kTY);
    const size_t smem = sizeof(float) * 2 * kTY * (size_t)W;
    SIDE_REQUIRE(smem <= 200 * 1024, "side_concat_volume_fwd: W=%d too wide", W);
    const int use_tma = ((W & 3) == 0) && aligned16(L) && aligned16(R) && aligned16(vol);
    int rc = set_smem_attr((const void *)concat_volume_fwd_kernel, smem);
    if (rc) return rc;
    const long long blocks = (long long)B * C * ytiles;
    SIDE_REQUIRE(blocks < (1ll << 31), "side_concat_volume_fwd: grid too large");
    concat_volume_fwd_kernel<<<(unsigned)blocks, kVolBlock, smem, (cudaStream_t)stream>>>(L, R, vol, C, H, W, D, ytiles,
                                                                                         use_tma);
    SIDE_LAUNCH_CHECK("concat_volume_fwd_kernel");
    return SIDE_OK;
}

extern "C" int side_concat_volume_bwd(const float *gvol, float *gL, float *gR, int B, int C, int H, int W, int D,
                                      void *stream)
{
    SIDE_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0, "side_concat_volume_bwd: bad shape");
    SIDE_REQUIRE((long long)B * C <= 65535, "side_concat_volume_bwd: B*C > 65535");
    SIDE_REQUIRE_DEV(gvol); SIDE_REQUIRE_DEV(gL); SIDE_REQUIRE_DEV(gR);
    dim3 grid(ceil_div((long long)H * W, kVolBlock), B * C);
    concat_volume_bwd_kernel<<<grid, kVolBlock, 0, (cudaStream_t)stream>>>(gvol, gL, gR, C, H, W, D);
    SIDE_LAUNCH_CHECK("concat_volume_bwd_kernel");
    return SIDE_OK;
}

extern "C" int side_gwc_volume_fwd(const float *L, const float *R, float *vol, int B, int C, int H, int W, int D, int G,
                                   void *stream)
{
    SIDE_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0 && G > 0 && C % G == 0, "side_gwc_volume_fwd: bad shape");
    SIDE_REQUIRE_DEV(L); SIDE_REQUIRE_DEV(R); SIDE_REQUIRE_DEV(vol);
    const int cpg = C / G;
    const size_t smem = sizeof(float) * 2 * (size_t)cpg * W;
    SIDE_REQUIRE(smem <= 200 * 1024, "side_gwc_volume_fwd: (C/G)*W too large for shared memory");
    const int use_tma = ((W & 3) == 0) && aligned16(L) && aligned16(R) && cpg <= 64;
    const long long blocks = (long long)B * G * H;
    SIDE_REQUIRE(blocks < (1ll << 31), "side_gwc_volume_fwd: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (use_tma && (D & 3) == 0 && (cpg == 4 || cpg == 8 || cpg == 16)) {
        const size_t smem_t = sizeof(float) * ((size_t)cpg * W + (size_t)cpg * (W + D + 4));
        if (smem_t <= 200 * 1024) {
#define GWC_TILED(CPGT)                                                                                         \
    do {                                                                                                        \
        if ((rc = set_smem_attr((const void *)gwc_volume_fwd_tiled_kernel<CPGT>, smem_t))) return rc;            \
        gwc_volume_fwd_tiled_kernel<CPGT><<<(unsigned)blocks, kVolBlock, smem_t, st>>>(L, R, vol, C, H, W, D, G); \
    } while (0)
            if (cpg == 4) GWC_TILED(4);
            else if (cpg == 8) GWC_TILED(8);
            else GWC_TILED(16);
#undef GWC_TILED
            SIDE_LAUNCH_CHECK("gwc_volume_fwd_tiled_kernel");
            return SIDE_OK;
        }
    }
#define GWC_LAUNCH(CPGT)                                                                                         \
    do {                                                                                                         \
        if ((rc = set_smem_attr((const void *)gwc_volume_fwd_kernel<CPGT>, smem))) return rc;                    \
        gwc_volume_fwd_kernel<CPGT><<<(unsigned)blocks, kVolBlock, smem, st>>>(L, R, vol, C, H, W, D, G, cpg, use_tma); \
    } while (0)
    switch (cpg) {
        case 4: GWC_LAUNCH(4); break;
        case 8: GWC_LAUNCH(8); break;
        case 16: GWC_LAUNCH(16); break;
        default: GWC_LAUNCH(0); break;
    }
#undef GWC_LAUNCH
    SIDE_LAUNCH_CHECK("gwc_volume_fwd_kernel");
    return SIDE_OK;
}

extern "C" int side_gwc_volume_bwd(const float *L, const float *R, const float *gvol, float *gL, float *gR, int B, int C,
                                   int H, int W, int D, int G, void *stream)
{
    SIDE_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && D > 0 && G > 0 && C % G == 0, "side_gwc_volume_bwd: bad shape");
    SIDE_REQUIRE_DEV(L); SIDE_REQUIRE_DEV(R); SIDE_REQUIRE_DEV(gvol); SIDE_REQUIRE_DEV(gL); SIDE_REQUIRE_DEV(gR);
    const int cpg = C / G;
    const size_t smem = sizeof(float) * 2 * (size_t)cpg * W;
    SIDE_REQUIRE(smem <= 200 * 1024, "side_gwc_volume_bwd: (C/G)*W too large for shared memory");
    int rc = set_smem_attr((const void *)gwc_volume_bwd_kernel, smem);
    if (rc) return rc;
    const long long blocks = (long long)B * G * H;
    SIDE_REQUIRE(blocks < (1ll << 31), "side_gwc_volume_bwd: grid too large");
    gwc_volume_bwd_kernel<<<(unsigned)blocks, kVolBlock, smem, (cudaStream_t)stream>>>(L, R, gvol, gL, gR, C, H, W, D, G,
                                                                                      cpg);
    SIDE_LAUNCH_CHECK("gwc_volume_bwd_kernel");
    return SIDE_OK;
}
