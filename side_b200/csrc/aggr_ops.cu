// aggr_ops.cu -- the small HBM-bound steps between the tensor-core convolutions of the aggregation network
// (cost_volume.forward, stereo_network_old.py:205-227), all on channels-last (NDHWC) activations:
//   layout change NCDHW -> NDHWC with the tf32 hi/lo split          (input of dres0)
//   structure-aware gate  cost = isp * cost  (:207-210)              fused with the split for dres1
//   MaxPool3d((1,2,2))    (:213, :218)                               fused with the split for dres2 / classify
//   classify's last Conv3d(64 -> 1) (:170)                           warp-per-voxel SIMT dot products
// Every kernel reads its input once and writes each output once (vectorised 16-byte accesses).
#include "tc_common.cuh"

namespace side {

int launch_cl_split_tile_f16(const float *x, float *full, void *hi, void *lo, int N, int C, long long S, int Cpad, cudaStream_t st);

__device__ __forceinline__ void split4(const float4 v, float4 &h, float4 &l)
{
    h.x = tf32_hi(v.x); l.x = v.x - h.x;
    h.y = tf32_hi(v.y); l.y = v.y - h.y;
    h.z = tf32_hi(v.z); l.z = v.z - h.z;
    h.w = tf32_hi(v.w); l.w = v.w - h.w;
}

// store the split of 4 consecutive channels at float4-index i: tf32 pairs (fp32 arrays) or fp16 pairs (half arrays)
template <bool F16>
__device__ __forceinline__ void store_split4(float4 *__restrict__ hi, float4 *__restrict__ lo, long long i, const float4 v)
{
    if (F16) {
        uint32_t h0, l0, h1, l1;
        f16_split2(v.x, v.y, h0, l0);
        f16_split2(v.z, v.w, h1, l1);
        reinterpret_cast<uint2 *>(hi)[i] = make_uint2(h0, h1);
        reinterpret_cast<uint2 *>(lo)[i] = make_uint2(l0, l1);
    } else {
        float4 h, l;
        split4(v, h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

// in [N, C, S] -> hi, lo [N, S, C]; optional scale[n, p / per_d] (the cosine gate of the slice the voxel belongs to)
template <bool F16>
__global__ void ncdhw_to_cl_split_kernel(const float *__restrict__ in, const float *__restrict__ scale, float *__restrict__ full,
                                         float *__restrict__ hi, float *__restrict__ lo, int C, long long S, int D, int per_d,
                                         int Cpad, uint32_t *rs)
{
    float amax = 0.f;                    // fp16 range guard (tc_common.cuh)
    // outputs have Cpad >= C channels per voxel, the extra ones zero (fp16 k-blocks are 64 channels wide)
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float *ip = in + (size_t)n * C * S;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long long p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < S) ? __ldg(ip + (size_t)c * S + p) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long p = p0 + i;
        const int c = c0 + threadIdx.x;
        if (p < S && c < Cpad) {
            float v = c < C ? tile[threadIdx.x][i] : 0.f;
            if (scale) v = __fmul_rn(v, __ldg(scale + (size_t)n * D + (int)(p / per_d)));
            const size_t o = ((size_t)n * S + p) * Cpad + c;
            if (F16) {
                __half h, l;
                amax = fmaxf(amax, fabsf(v));
                f16_split(v, h, l);
                reinterpret_cast<__half *>(hi)[o] = h;
                reinterpret_cast<__half *>(lo)[o] = l;
            } else {
                const float h = tf32_hi(v);
                hi[o] = h;
                lo[o] = v - h;
            }
            if (full) full[o] = v;
        }
    }
    if (F16 && rs) range_commit(rs, amax);
}

// fp16 variant with 64-channel tiles: a lane owns two adjacent channels, so every warp store is 128 contiguous bytes of
// one voxel's channels-last row (the generic kernel above would issue 2-byte stores)
__global__ void __launch_bounds__(256) ncdhw_to_cl_split_f16_kernel(const float *__restrict__ in, const float *__restrict__ scale,
                                                                   float *__restrict__ full, uint32_t *__restrict__ hi,
                                                                   uint32_t *__restrict__ lo, int C, long long S, int D, int per_d,
                                                                   int Cpad, uint32_t *rs)
{
    __shared__ float tile[64][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 64;
    const float *ip = in + (size_t)n * C * S;
    for (int i = threadIdx.y; i < 64; i += blockDim.y) {
        const int c = c0 + i;
        const long long p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < S) ? __ldg(ip + (size_t)c * S + p) : 0.f;
    }
    __syncthreads();
    const int c = c0 + 2 * threadIdx.x;
    float amax = 0.f;
    for (int i = threadIdx.y; i < 32 && c < Cpad; i += blockDim.y) {
        const long long p = p0 + i;
        if (p >= S) continue;
        float a = tile[2 * threadIdx.x][i], b = tile[2 * threadIdx.x + 1][i];
        if (scale) {
            const float g = __ldg(scale + (size_t)n * D + (int)(p / per_d));
            a = __fmul_rn(a, g);
            b = __fmul_rn(b, g);
        }
        const size_t o = (((size_t)n * S + p) * Cpad + c) >> 1;      // index of the channel PAIR
        uint32_t h, l;
        amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
        f16_split2(a, b, h, l);
        hi[o] = h;
        lo[o] = l;
        if (full) reinterpret_cast<float2 *>(full)[o] = make_float2(a, b);
    }
    if (rs) range_commit(rs, amax);
}

__global__ void tf32_split_kernel(const float4 *__restrict__ x, float4 *__restrict__ hi, float4 *__restrict__ lo, long long n4)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 h, l;
        split4(__ldg(x + i), h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

// fp32 -> fp16 (hi, lo * 2^11) pairs of a tensor that is already in the consumer's layout
__global__ void f16_split_kernel(const float4 *__restrict__ x, uint2 *__restrict__ hi, uint2 *__restrict__ lo, long long n4,
                                 uint32_t *rs)
{
    float amax = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        uint2 h, l;
        f16_split2(v.x, v.y, h.x, l.x);
        f16_split2(v.z, v.w, h.y, l.y);
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        hi[i] = h;
        lo[i] = l;
    }
    if (rs) range_commit(rs, amax);
}

// y [N, D, H, W, C] * gate [N, D, W, C] (broadcast over H) -> hi, lo
template <bool F16>
__global__ void gate_mul_split_kernel(const float4 *__restrict__ y, const float4 *__restrict__ gate, float4 *__restrict__ hi,
                                      float4 *__restrict__ lo, long long n4, int H, int WC4, uint32_t *rs)
{
    float amax = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / WC4;             // (n, d, h)
        const int wc = (int)(i - row * WC4);
        const long long nd = row / H;
        const float4 v = __ldg(y + i), g = __ldg(gate + nd * WC4 + wc);
        const float4 o = make_float4(v.x * g.x, v.y * g.y, v.z * g.z, v.w * g.w);
        if (F16) amax = fmaxf(fmaxf(amax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
        store_split4<F16>(hi, lo, i, o);
    }
    if (F16 && rs) range_commit(rs, amax);
}

// x [ND, H, W, C] -> max over 2x2 (h, w) windows -> [ND, H/2, W/2, C]; y and/or (hi, lo)
template <bool F16>
__global__ void maxpool_hw2_cl_kernel(const float4 *__restrict__ x, float4 *__restrict__ y, float4 *__restrict__ hi,
                                      float4 *__restrict__ lo, long long n4out, int H, int W, int C4, uint32_t *rs)
{
    const int Ho = H / 2, Wo = W / 2;
    float amax = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4out; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        long long r = i / C4;
        const int wo = (int)(r % Wo);
        r /= Wo;
        const int ho = (int)(r % Ho);
        const long long nd = r / Ho;
        const float4 *p = x + ((nd * H + 2 * ho) * W + 2 * wo) * C4 + c;
        const float4 a = __ldg(p), b = __ldg(p + C4), cc = __ldg(p + (size_t)W * C4), d = __ldg(p + (size_t)W * C4 + C4);
        float4 m;
        m.x = fmaxf(fmaxf(a.x, b.x), fmaxf(cc.x, d.x));
        m.y = fmaxf(fmaxf(a.y, b.y), fmaxf(cc.y, d.y));
        m.z = fmaxf(fmaxf(a.z, b.z), fmaxf(cc.z, d.z));
        m.w = fmaxf(fmaxf(a.w, b.w), fmaxf(cc.w, d.w));
        if (y) y[i] = m;
        if (hi) {
            if (F16) amax = fmaxf(fmaxf(amax, fmaxf(fabsf(m.x), fabsf(m.y))), fmaxf(fabsf(m.z), fabsf(m.w)));
            store_split4<F16>(hi, lo, i, m);
        }
    }
    if (F16 && rs && hi) range_commit(rs, amax);
}

// x [N, D, H, W, C], w [1, C, 3, 3, 3] -> out [N, D, H, W]   (Conv3d(C -> 1, 3, padding 1, bias=False)).
// C / 4 lanes per voxel (16 for C = 64, two voxels per warp): one 16-byte activation load and one 16-byte shared weight load
// per tap and lane, all 27 taps unrolled so the loads of a voxel are in flight together, then a butterfly over the lanes.
template <int LPV>
__global__ void __launch_bounds__(256) conv3d_c1_cl_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                           float *__restrict__ out, long long nvox, int D, int H, int W, int C)
{
    extern __shared__ __align__(16) float ws[];   // [27][C]
    for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) {
        const int tap = i / C, c = i - tap * C;
        ws[i] = __ldg(w + (size_t)c * 27 + tap);
    }
    __syncthreads();
    constexpr int VPW = 32 / LPV;                            // voxels per warp
    const int lane = threadIdx.x & 31, sub = lane % LPV;     // channel quad(s) of this lane: sub, sub + LPV, ...
    const long long w0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * VPW;
    const long long wstride = (((long long)gridDim.x * blockDim.x) >> 5) * VPW;
    const int C4 = C / 4;
    for (long long vw = w0; vw < nvox; vw += wstride) {      // warp-uniform loop: every lane reaches the shuffles
        const long long v = vw + lane / LPV;
        const bool live = v < nvox;
        const long long vv = live ? v : nvox - 1;
        const int wq = (int)(vv % W);
        long long r = vv / W;
        const int hq = (int)(r % H);
        r /= H;
        const int dq = (int)(r % D);
        const long long n = r / D;
        float acc = 0.f;
#pragma unroll
        for (int tap = 0; tap < 27; ++tap) {
            const int dd = dq + tap / 9 - 1, hh = hq + (tap / 3) % 3 - 1, ww = wq + tap % 3 - 1;
            if (dd < 0 || dd >= D || hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
            const float4 *xp = reinterpret_cast<const float4 *>(x + ((((size_t)n * D + dd) * H + hh) * W + ww) * C);
            const float4 *wp = reinterpret_cast<const float4 *>(ws + tap * C);
            for (int c = sub; c < C4; c += LPV) {
                const float4 a = __ldg(xp + c), b = wp[c];
                acc = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
            }
        }
#pragma unroll
        for (int o = LPV / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (sub == 0 && live) out[v] = acc;
    }
}

// The same layer for small per-sample volumes (the classifier of the aggregation network: 16 x 4 x 4 voxels per RoI, C = 64).
// The gather form above reads every voxel's channels 27 times (2.8 GB of L2 traffic for 105 MB of activations: 0.75 ms per 1600
// RoIs).  Here a CTA owns one sample and a thread one voxel: it reads its C = 64 channels ONCE into registers, projects them on
// all 27 tap filters (P[tap][voxel], weights broadcast from shared memory), and after one barrier sums the 27 projections of its
// neighbours:  out[v] = sum_tap P[tap][v + tap].  Same products, summed per tap first (<= 1e-6 relative to the gather form).
template <int C>
__global__ void __launch_bounds__(512) conv3d_c1_roi_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                            float *__restrict__ out, int D, int H, int W)
{
    extern __shared__ __align__(16) float sm[];       // ws [27][C], then P [27][nv + 1]
    const int nv = D * H * W, tid = threadIdx.x;
    float *ws = sm, *P = sm + 27 * C;
    for (int i = tid; i < 27 * C; i += blockDim.x) {
        const int tap = i / C, c = i - tap * C;
        ws[i] = __ldg(w + (size_t)c * 27 + tap);
    }
    float xr[C];
    if (tid < nv) {
        const float4 *xp = reinterpret_cast<const float4 *>(x + ((size_t)blockIdx.x * nv + tid) * C);
#pragma unroll
        for (int c = 0; c < C / 4; ++c) {
            const float4 v = __ldg(xp + c);
            xr[4 * c] = v.x; xr[4 * c + 1] = v.y; xr[4 * c + 2] = v.z; xr[4 * c + 3] = v.w;
        }
    }
    __syncthreads();
    if (tid < nv) {
#pragma unroll 1
        for (int tap = 0; tap < 27; ++tap) {
            const float4 *wp = reinterpret_cast<const float4 *>(ws + tap * C);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int c = 0; c < C / 4; ++c) {
                const float4 b = wp[c];
                a0 = fmaf(xr[4 * c], b.x, a0); a1 = fmaf(xr[4 * c + 1], b.y, a1);
                a2 = fmaf(xr[4 * c + 2], b.z, a2); a3 = fmaf(xr[4 * c + 3], b.w, a3);
            }
            P[tap * (nv + 1) + tid] = (a0 + a1) + (a2 + a3);
        }
    }
    __syncthreads();
    if (tid < nv) {
        const int wq = tid % W, hq = (tid / W) % H, dq = tid / (W * H);
        float acc = 0.f;
#pragma unroll
        for (int tap = 0; tap < 27; ++tap) {
            const int dd = dq + tap / 9 - 1, hh = hq + (tap / 3) % 3 - 1, ww = wq + tap % 3 - 1;
            if (dd < 0 || dd >= D || hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
            acc += P[tap * (nv + 1) + (dd * H + hh) * W + ww];
        }
        out[(size_t)blockIdx.x * nv + tid] = acc;
    }
}

}  // namespace side

using namespace side;

static inline unsigned ew_grid(long long n, int block) { return (unsigned)std::min<long long>((n + block - 1) / block, 148 * 16); }

static int ncdhw_to_cl_split_impl(const float *x, const float *scale, float *full, float *hi, float *lo, int N, int C,
                                  long long S, int D, void *stream, bool f16, int Cpad);
extern "C" int side_ncdhw_to_cl_split(const float *x, const float *scale, float *full, float *hi, float *lo, int N, int C,
                                      long long S, int D, void *stream)
{
    return ncdhw_to_cl_split_impl(x, scale, full, hi, lo, N, C, S, D, stream, false, C);
}
/* the _f16 variants write the fp16 operand pairs of side_conv3d_tc_fwd_f16 (hi, lo: half arrays of the same element count) */
extern "C" int side_ncdhw_to_cl_split_f16(const float *x, const float *scale, float *full, void *hi, void *lo, int N, int C,
                                          long long S, int D, int Cpad, void *stream)
{
    return ncdhw_to_cl_split_impl(x, scale, full, reinterpret_cast<float *>(hi), reinterpret_cast<float *>(lo), N, C, S, D, stream,
                                  true, Cpad);
}
static int ncdhw_to_cl_split_impl(const float *x, const float *scale, float *full, float *hi, float *lo, int N, int C,
                                  long long S, int D, void *stream, bool f16, int Cpad)
{
    if (Cpad <= 0) Cpad = C;
    SIDE_REQUIRE(Cpad >= C, "side_ncdhw_to_cl_split: Cpad < C");
    SIDE_REQUIRE(N >= 0 && C > 0 && S > 0, "side_ncdhw_to_cl_split: bad shape");
    SIDE_REQUIRE(scale == nullptr || (D > 0 && S % D == 0), "side_ncdhw_to_cl_split: scale needs D | S");
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE(N <= 65535 && (C + 31) / 32 <= 65535, "side_ncdhw_to_cl_split: grid too large");
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(hi); SIDE_REQUIRE_DEV(lo);
    dim3 g((unsigned)((S + 31) / 32), (unsigned)((Cpad + 31) / 32), (unsigned)N), b(32, 8);
    if (scale) SIDE_REQUIRE_DEV(scale);
    if (full) SIDE_REQUIRE_DEV(full);
    const int per_d = scale ? (int)(S / D) : 1;
    if (f16 && !scale && S % 4 == 0 && Cpad % 8 == 0 && N <= 65535 &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(full) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0)
        return launch_cl_split_tile_f16(x, full, hi, lo, N, C, S, Cpad, (cudaStream_t)stream);     // 16-byte loads, 64 x 128 tiles (upsample.cu)
    if (f16 && Cpad % 2 == 0) {
        dim3 g2((unsigned)((S + 31) / 32), (unsigned)((Cpad + 63) / 64), (unsigned)N);
        ncdhw_to_cl_split_f16_kernel<<<g2, b, 0, (cudaStream_t)stream>>>(x, scale, full, reinterpret_cast<uint32_t *>(hi),
                                                                         reinterpret_cast<uint32_t *>(lo), C, S, D, per_d, Cpad,
                                                                         range_slot_next());
    } else if (f16) ncdhw_to_cl_split_kernel<true><<<g, b, 0, (cudaStream_t)stream>>>(x, scale, full, hi, lo, C, S, D, per_d, Cpad,
                                                                                      range_slot_next());
    else ncdhw_to_cl_split_kernel<false><<<g, b, 0, (cudaStream_t)stream>>>(x, scale, full, hi, lo, C, S, D, per_d, Cpad, nullptr);
    SIDE_LAUNCH_CHECK("ncdhw_to_cl_split_kernel");
    return SIDE_OK;
}

extern "C" int side_tf32_split(const float *x, float *hi, float *lo, long long n, void *stream)
{
    SIDE_REQUIRE(n >= 0 && n % 4 == 0, "side_tf32_split: element count must be a multiple of 4");
    if (n == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(hi); SIDE_REQUIRE_DEV(lo);
    tf32_split_kernel<<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(x), reinterpret_cast<float4 *>(hi), reinterpret_cast<float4 *>(lo), n / 4);
    SIDE_LAUNCH_CHECK("tf32_split_kernel");
    return SIDE_OK;
}

extern "C" int side_f16_split(const float *x, void *hi, void *lo, long long n, void *stream)
{
    SIDE_REQUIRE(n >= 0 && n % 4 == 0, "side_f16_split: element count must be a multiple of 4");
    if (n == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(hi); SIDE_REQUIRE_DEV(lo);
    f16_split_kernel<<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(x),
                                                                           reinterpret_cast<uint2 *>(hi), reinterpret_cast<uint2 *>(lo),
                                                                           n / 4, range_slot_next());
    SIDE_LAUNCH_CHECK("f16_split_kernel");
    return SIDE_OK;
}

static int gate_mul_split_impl(const float *y, const float *gate, float *hi, float *lo, int N, int D, int H, int W, int C,
                               void *stream, bool f16);
extern "C" int side_gate_mul_split(const float *y, const float *gate, float *hi, float *lo, int N, int D, int H, int W, int C,
                                   void *stream)
{
    return gate_mul_split_impl(y, gate, hi, lo, N, D, H, W, C, stream, false);
}
extern "C" int side_gate_mul_split_f16(const float *y, const float *gate, void *hi, void *lo, int N, int D, int H, int W, int C,
                                       void *stream)
{
    return gate_mul_split_impl(y, gate, reinterpret_cast<float *>(hi), reinterpret_cast<float *>(lo), N, D, H, W, C, stream, true);
}
static int gate_mul_split_impl(const float *y, const float *gate, float *hi, float *lo, int N, int D, int H, int W, int C,
                               void *stream, bool f16)
{
    SIDE_REQUIRE(N >= 0 && D > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "side_gate_mul_split: bad shape (C %% 4 == 0)");
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(y); SIDE_REQUIRE_DEV(gate); SIDE_REQUIRE_DEV(hi); SIDE_REQUIRE_DEV(lo);
    const long long n4 = (long long)N * D * H * W * C / 4;
    if (f16)
        gate_mul_split_kernel<true><<<ew_grid(n4, 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4 *>(y), reinterpret_cast<const float4 *>(gate), reinterpret_cast<float4 *>(hi),
            reinterpret_cast<float4 *>(lo), n4, H, W * C / 4, range_slot_next());
    else
        gate_mul_split_kernel<false><<<ew_grid(n4, 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4 *>(y), reinterpret_cast<const float4 *>(gate), reinterpret_cast<float4 *>(hi),
            reinterpret_cast<float4 *>(lo), n4, H, W * C / 4, nullptr);
    SIDE_LAUNCH_CHECK("gate_mul_split_kernel");
    return SIDE_OK;
}

static int maxpool_hw2_cl_impl(const float *x, float *y, float *hi, float *lo, int N, int D, int H, int W, int C, void *stream,
                               bool f16);
extern "C" int side_maxpool_hw2_cl(const float *x, float *y, float *hi, float *lo, int N, int D, int H, int W, int C,
                                   void *stream)
{
    return maxpool_hw2_cl_impl(x, y, hi, lo, N, D, H, W, C, stream, false);
}
extern "C" int side_maxpool_hw2_cl_f16(const float *x, float *y, void *hi, void *lo, int N, int D, int H, int W, int C,
                                       void *stream)
{
    return maxpool_hw2_cl_impl(x, y, reinterpret_cast<float *>(hi), reinterpret_cast<float *>(lo), N, D, H, W, C, stream, true);
}
static int maxpool_hw2_cl_impl(const float *x, float *y, float *hi, float *lo, int N, int D, int H, int W, int C, void *stream,
                               bool f16)
{
    SIDE_REQUIRE(N >= 0 && D > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 4 == 0,
                 "side_maxpool_hw2_cl: bad shape (even H, W; C %% 4 == 0)");
    SIDE_REQUIRE(y || (hi && lo), "side_maxpool_hw2_cl: no output requested");
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(x);
    if (y) SIDE_REQUIRE_DEV(y);
    if (hi) { SIDE_REQUIRE_DEV(hi); SIDE_REQUIRE_DEV(lo); }
    const long long n4 = (long long)N * D * (H / 2) * (W / 2) * C / 4;
    if (f16)
        maxpool_hw2_cl_kernel<true><<<ew_grid(n4, 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4 *>(x), reinterpret_cast<float4 *>(y), reinterpret_cast<float4 *>(hi),
            reinterpret_cast<float4 *>(lo), n4, H, W, C / 4, range_slot_next());
    else
        maxpool_hw2_cl_kernel<false><<<ew_grid(n4, 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const float4 *>(x), reinterpret_cast<float4 *>(y), reinterpret_cast<float4 *>(hi),
            reinterpret_cast<float4 *>(lo), n4, H, W, C / 4, nullptr);
    SIDE_LAUNCH_CHECK("maxpool_hw2_cl_kernel");
    return SIDE_OK;
}

extern "C" int side_conv3d_c1_cl(const float *x, const float *w, float *out, int N, int D, int H, int W, int C, void *stream)
{
    SIDE_REQUIRE(N >= 0 && D > 0 && H > 0 && W > 0 && C > 0 && 27 * C * 4 <= 48 * 1024, "side_conv3d_c1_cl: bad shape");
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(x); SIDE_REQUIRE_DEV(w); SIDE_REQUIRE_DEV(out);
    const long long nvox = (long long)N * D * H * W;
    SIDE_REQUIRE(C % 4 == 0, "side_conv3d_c1_cl: C %% 4 == 0");
    const int nv = D * H * W;
    if (C == 64 && nv <= 512 && N <= (1 << 30) && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {       // one sample per CTA, one voxel per thread
        const size_t smem = sizeof(float) * (27 * 64 + 27 * (size_t)(nv + 1));
        int rc;
        if ((rc = set_smem_attr((const void *)conv3d_c1_roi_kernel<64>, smem))) return rc;
        conv3d_c1_roi_kernel<64><<<(unsigned)N, (unsigned)((nv + 31) / 32 * 32), smem, (cudaStream_t)stream>>>(x, w, out, D, H, W);
        SIDE_LAUNCH_CHECK("conv3d_c1_roi_kernel");
        return SIDE_OK;
    }
    if (C >= 128) conv3d_c1_cl_kernel<32><<<ew_grid(nvox * 32, 256), 256, 27 * C * sizeof(float), (cudaStream_t)stream>>>(x, w, out, nvox, D, H, W, C);
    else if (C >= 64) conv3d_c1_cl_kernel<16><<<ew_grid(nvox * 16, 256), 256, 27 * C * sizeof(float), (cudaStream_t)stream>>>(x, w, out, nvox, D, H, W, C);
    else conv3d_c1_cl_kernel<4><<<ew_grid(nvox * 4, 256), 256, 27 * C * sizeof(float), (cudaStream_t)stream>>>(x, w, out, nvox, D, H, W, C);
    SIDE_LAUNCH_CHECK("conv3d_c1_cl_kernel");
    return SIDE_OK;
}
