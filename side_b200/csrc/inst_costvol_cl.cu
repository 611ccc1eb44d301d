// inst_costvol_cl.cu -- instance depth cost volume emitted DIRECTLY in the format its consumer reads: gated, channels-last
// [N, D, 16, 16, 3C] and already split into the fp16 operand pairs (hi, lo * 2^11) of the tcgen05 convolution dres0.0
// (SURVEY.md section 8 rows A4, A5, A6-gate feeding F1).
//
// Reference behaviour replaced (stereo_network_old.py): get_proposal_shift (:34-133), the 2*D RoIAlign launches and 3*D slice
// copies that build cost[N, 3C, D, 16, 16] (:366-376), and the cosine gate cost *= x_cross[n, d] (:197-203).  Round 1 wrote
// that volume as fp32 NCDHW (inst_costvol_sep_kernel) and a second kernel re-read it to produce the channels-last pairs:
// 2 x 1.26 GB of HBM traffic per 800 RoIs for data that is consumed exactly once.  Here the volume crosses HBM once.
//
// Work decomposition (C = 32 channels per view, the reference's reduced_channel):
//   CTA  = one RoI at a time (persistent grid, RoIs handed out by an atomic counter), 16 warps = the 16 bin rows ph
//   lane = (voxel of a pair, channel pair): a half-warp covers the 32 channels of one bin with 8-byte accesses, so one
//       shared-memory instruction serves two bins, the fp16 conversion packs a lane's two channels directly and no
//       cross-lane exchange is needed anywhere
//   U[side][ph][x][c] = sum over the bin row's two y samples of (hy f[ylo][x][c] + ly f[yhi][x][c]) -- the y half of the
//       bilinear interpolation, identical for all D candidates of a RoI (they only shift the box along x) -- is built once per
//       group of slices for the window of columns the group's boxes touch (fp32, 160 KB of shared memory)
//   slice d: every warp evaluates its 16 bins x (L, R) from U (4 taps each), the three gate sums are reduced over the CTA,
//       and only THEN are the values scaled by x_cross, split into fp16 pairs and staged so that a quarter bin row =
//       4 voxels x 96 channels = 768 contiguous bytes of the destination leaves as ONE TMA bulk store per array
//       (cp.async.bulk shared -> global), double-buffered per warp.  No LSU store instruction touches the volume.
// Boxes too wide for the window (a slice's x range > 38 columns) are processed in 2 .. 16 column tiles with a statistics
// pass before the emitting pass (the gate needs the sums of the whole slice before its first value may be written).
// OC = 64 (SIDE_VOL_NO_DIFF): only the L and R planes are emitted, [N, D, 16, 16, 2C].  The L - R plane of the reference's
// cat(L, R, L - R) (stereo_network_old.py:374-376) is a linear function of the other two, so its consumer dres0.0 -- the only
// reader of the volume -- folds it into the weights instead (W_L + W_D, W_R - W_D: side_b200/networks/stereo_network.py):
// a third less volume to write and a 64-channel first convolution (one k-block per tap instead of one and a half).
// Numerics: the same sample positions, validity rules, weights and operation order as inst_costvol_sep_kernel (<= 1e-5
// relative to torchvision's RoIAlign order); the pair (hi, lo) carries 22 significand bits.
#include <cuda_fp16.h>

#include "tc_common.cuh"
#include "vol_common.cuh"

namespace side {

constexpr int kClC = 32;                          // channels per view
constexpr int kClThreads = 512;                   // 16 warps = 16 bin rows
constexpr int kClXW = 38;                         // window columns per view
constexpr int kClCells = kClXW + 2;               // + two zero cells (invalid samples, right border)
constexpr int kClMaxD = 64;
constexpr int kClSub = 16;                        // slices whose x-sample tables are held at once
constexpr int kClRowF = kClCells * kClC;          // floats per (view, bin row)
constexpr int kClSideF = 16 * kClRowF;            // floats per view
constexpr int kClChunkB = 4 * 96 * 2;             // bytes of one staged chunk per array: 4 voxels x 96 channels x fp16
constexpr int kClWarpStageB = 2 * 2 * kClChunkB;  // per warp: 2 buffers x (hi, lo)
constexpr size_t kClUBytes = sizeof(float) * 2 * kClSideF;
constexpr size_t kClStageBytes = 16 * kClWarpStageB;
constexpr size_t kClXtabBytes = sizeof(float4) * kClSub * 2 * 16;
constexpr size_t kClSmem = kClUBytes + kClStageBytes + kClXtabBytes;

struct ClParams {
    VolParams v;
    const float *nhwcL, *nhwcR;       // [B, H, W, 32]
    __half *hi, *lo;                  // [N, D, 16, 16, 96]
    int *counter;                     // work counter (zeroed by the host before the launch)
    int gate;                         // 1: values are multiplied by x_cross[n, d] before the split
    uint32_t *rs;                     // fp16 range-guard slot or NULL
};

__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// two floats -> one f16x2 word (a in the low half), saturating
__device__ __forceinline__ uint32_t cl_pack_h2(float a, float b)
{
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// One plane (L, R or L - R) of this lane's voxel, channels (2 cp, 2 cp + 1): split into (hi, lo') words and staged.  The two
// half-warps write two voxel rows 48 words apart: disjoint bank halves, conflict-free.
__device__ __forceinline__ void cl_stage_plane(float x0, float x1, uint32_t hi_addr, uint32_t lo_addr)
{
    const uint32_t hw = cl_pack_h2(x0, x1);
    const __half2 h2 = *reinterpret_cast<const __half2 *>(&hw);
    const uint32_t lw = cl_pack_h2((x0 - __low2float(h2)) * kF16LoScale, (x1 - __high2float(h2)) * kF16LoScale);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(hi_addr), "r"(hw) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(lo_addr), "r"(lw) : "memory");
}

// One chunk = NS steps of this warp's bin row (a step = 2 voxels, one per half-warp) -> staged -> two bulk stores of nbytes.
// l[], r[]: the steps' values of this lane's channel pair; buf: this warp's staging buffer for the chunk ([hi][lo]);
// slot0: staging slot of the lane's voxel in step 0 (step j uses slot0 + 2 j); vox: index of the chunk's first voxel.
template <int NS, int OC>
__device__ __forceinline__ void cl_emit_chunk(const float2 *l, const float2 *r, float g, unsigned char *buf, __half *ghi, __half *glo,
                                              size_t vox, uint32_t nbytes, int lane, int slot0)
{
    if (lane == 0) bulk_wait_read<1>();          // the store issued from this buffer two chunks ago has read it
    __syncwarp();
    // OC = 64: a voxel row is 32 words, the two half-warps fall on the same banks (2-way conflict on these stores)
    const uint32_t base = smem_u32(buf) + (uint32_t)((slot0 * (OC / 2) + (lane & 15)) * 4);
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const uint32_t a = base + (uint32_t)(j * 2 * (OC * 2));
        const float2 lv = l[j], rv = r[j];
        cl_stage_plane(__fmul_rn(lv.x, g), __fmul_rn(lv.y, g), a, a + kClChunkB);
        cl_stage_plane(__fmul_rn(rv.x, g), __fmul_rn(rv.y, g), a + 64, a + 64 + kClChunkB);
        if (OC == 96)
            cl_stage_plane(__fmul_rn(__fsub_rn(lv.x, rv.x), g), __fmul_rn(__fsub_rn(lv.y, rv.y), g), a + 128, a + 128 + kClChunkB);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        bulk_s2g(ghi + vox * OC, buf, nbytes);
        bulk_s2g(glo + vox * OC, buf + kClChunkB, nbytes);
        bulk_commit();
    }
}

// 4-tap value of one bin for this lane's channel pair: e = {byte offset of sample 0's first cell, 0.25 * l0, offset 1, 0.25 * l1}
__device__ __forceinline__ float2 cl_bin(uint32_t urow, const float4 e)
{
    const uint32_t a0 = urow + __float_as_uint(e.x), a1 = urow + __float_as_uint(e.z);
    float2 p0, p1, p2, p3;
    // volatile: these read U, which other threads rewrite between barriers -- the compiler must not move them across one
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(p0.x), "=f"(p0.y) : "r"(a0));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+128];" : "=f"(p1.x), "=f"(p1.y) : "r"(a0));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(p2.x), "=f"(p2.y) : "r"(a1));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+128];" : "=f"(p3.x), "=f"(p3.y) : "r"(a1));
    const float w0 = e.y, w1 = e.w, wa0 = 0.25f - w0, wa1 = 0.25f - w1;
    return make_float2(fmaf(w1, p3.x, fmaf(wa1, p2.x, fmaf(w0, p1.x, wa0 * p0.x))),
                       fmaf(w1, p3.y, fmaf(wa1, p2.y, fmaf(w0, p1.y, wa0 * p0.y))));
}

// x_cross of a slice from 16 per-warp partial sums {sum L^2, sum R^2, sum L R}: fixed reduction order, every warp computes it
__device__ __forceinline__ float cl_gate(const float (*part)[4], int lane)
{
    float a = part[lane & 15][lane >> 4];            // lanes 0-15: sum L^2 partials, lanes 16-31: sum R^2 partials
    float c = part[lane & 15][2];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    const float t0 = __shfl_sync(0xffffffffu, a, 0), t1 = __shfl_sync(0xffffffffu, a, 16), t2 = __shfl_sync(0xffffffffu, c, 0);
    return __fdiv_rn(t2, fmaxf(__fmul_rn(__fsqrt_rn(t0), __fsqrt_rn(t1)), 0.01f));
}

template <int OC>
__global__ void __launch_bounds__(kClThreads, 1) inst_costvol_cl_kernel(ClParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *U = reinterpret_cast<float *>(smem_raw);                                   // [2][16][kClCells][32]
    unsigned char *stage = smem_raw + kClUBytes;                                       // [16 warps][2 buffers][hi | lo]
    float4 (*xtab)[2][16] = reinterpret_cast<float4 (*)[2][16]>(smem_raw + kClUBytes + kClStageBytes);   // [slice][view][bin column]
    __shared__ AxisTap ytab[32];
    __shared__ float4 geo[kClMaxD];                  // lx1, bin_w(left), rx1, bin_w(right) per slice
    __shared__ float part[2][16][4];                 // per-warp gate sums of the slice in flight (double-buffered)
    __shared__ float gtab[kClMaxD];                  // tiled mode: gate scalar per slice
    __shared__ int s_n, g_d1, g_win[2], g_wr[2];

    const VolParams &v = p.v;
    const int D = v.D, W = v.W, N = v.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hw = lane >> 4, cp = lane & 15;       // voxel of the pair, channel pair
    unsigned char *wstage = stage + warp * kClWarpStageB;
    float amax = 0.f;                                // max |L|, |R| this thread produced (range guard; |gate| <= 1)
    int cc = 0;                                      // chunks this warp has emitted so far: buffer = cc & 1, strictly alternating

    for (;;) {
        if (lane == 0) bulk_wait_read<0>();          // tiled mode reuses the staging area of ALL warps for its statistics
        __syncthreads();
        if (tid == 0) s_n = atomicAdd(p.counter, 1);
        __syncthreads();
        const int n = s_n;
        if (n >= N) break;
        __half *ghi = p.hi + (size_t)n * D * 256 * OC, *glo = p.lo + (size_t)n * D * 256 * OC;

        if (v.valid && !v.valid[n]) {                // dropped row of the fixed-shape RoI set: an all-zero volume, depth_bin = 0
            const int total = D * 256 * OC * 2 / 16;
            uint4 *a = reinterpret_cast<uint4 *>(ghi), *b = reinterpret_cast<uint4 *>(glo);
            for (int i = tid; i < total; i += kClThreads) {
                __stcs(a + i, make_uint4(0u, 0u, 0u, 0u));
                __stcs(b + i, make_uint4(0u, 0u, 0u, 0u));
            }
            for (int d = tid; d < D; d += kClThreads) {
                v.depth_bin[(size_t)n * D + d] = 0.f;
                if (v.xcross) v.xcross[(size_t)n * D + d] = 0.f;
            }
            continue;
        }

        const float *lb = v.left + (size_t)n * 5, *rb = v.right + (size_t)n * 5;
        const int b = min(max((int)lb[0], 0), v.B - 1);
        const float fb = v.fb[b];
        for (int d = tid; d < D; d += kClThreads) {
            float dbin, lx1, lx2, rx1, rx2, y1, y2;
            proposal_for(lb, rb, fb, d, D, v.x_clamp, dbin, lx1, lx2, rx1, rx2, y1, y2);
            const float rwl = fmaxf(__fsub_rn(lx2, lx1), 1.0f), rwr = fmaxf(__fsub_rn(rx2, rx1), 1.0f);
            geo[d] = make_float4(lx1, __fmul_rn(rwl, 0.0625f), rx1, __fmul_rn(rwr, 0.0625f));   // == rw / 16 exactly
            v.depth_bin[(size_t)n * D + d] = dbin;
        }
        if (tid < 32) {
            float dbin, lx1, lx2, rx1, rx2, y1, y2;
            proposal_for(lb, rb, fb, 0, D, v.x_clamp, dbin, lx1, lx2, rx1, rx2, y1, y2);
            const float rh = fmaxf(__fsub_rn(y2, y1), 1.0f);
            ytab[tid] = to_tap(axis_sample(y1, __fmul_rn(rh, 0.0625f), tid >> 1, tid & 1, v.H), W * kClC);
        }
        __syncthreads();

        // column tiles: the smallest power of two T such that every slice's x range of every tile fits the window
        int T = 1;
        for (; T < 16; T <<= 1) {
            const int S = 32 / T;
            bool ok = true;
            for (int i = tid; i < D * 2 * T; i += kClThreads) {
                const int d = i / (2 * T), r2 = i - d * 2 * T, side = r2 / T, t = r2 - side * T;
                const float4 g4 = geo[d];
                int c0, c1;
                sep_cells(side ? g4.z : g4.x, side ? g4.w : g4.y, t * S, t * S + S - 1, W, c0, c1);
                ok = ok && (c1 - c0 + 1 <= kClXW);
            }
            if (__syncthreads_and(ok)) break;
        }
        const int S = 32 / T, PWT = 16 / T;          // x-samples / bin columns per tile
        const size_t img = (size_t)b * v.H * W * kClC;
        const uint32_t urow = smem_u32(U + warp * kClRowF + 2 * cp);      // this lane's channel pair of bin row `warp`, left view

        // one pass over (tile, slice group, slice).  phase 0: single pass (T == 1: values kept in registers across the gate
        // reduction);  phase 1: statistics only (tiled);  phase 2: emit with the gate scalars of phase 1 (tiled)
        float (*partD)[16][4] = reinterpret_cast<float (*)[16][4]>(stage);            // tiled mode: sums per (slice, warp)
        if (T > 1) {
            for (int i = tid; i < D * 16 * 4; i += kClThreads) reinterpret_cast<float *>(stage)[i] = 0.f;
            __syncthreads();
        }
        for (int phase = (T == 1 ? 0 : 1); phase <= (T == 1 ? 0 : 2); ++phase) {
            for (int t = 0; t < T; ++t) {
                int d0 = 0;
                while (d0 < D) {
                    // ---- slice group [d0, d1]: grow while both views' windows stay within kClXW columns ----
                    if (warp == 0) {
                        const int side = lane & 1;
                        int lo, hi2, dd = d0;
                        {
                            const float4 g4 = geo[d0];
                            sep_cells(side ? g4.z : g4.x, side ? g4.w : g4.y, t * S, t * S + S - 1, W, lo, hi2);
                        }
                        while (dd + 1 < D) {
                            const float4 g4 = geo[dd + 1];
                            int c0, c1;
                            sep_cells(side ? g4.z : g4.x, side ? g4.w : g4.y, t * S, t * S + S - 1, W, c0, c1);
                            const int nlo = min(lo, c0), nhi = max(hi2, c1);
                            if (!__all_sync(0xffffffffu, nhi - nlo + 1 <= kClXW)) break;
                            lo = nlo; hi2 = nhi; ++dd;
                        }
                        if (lane < 2) { g_win[side] = lo; g_wr[side] = hi2 - lo + 1; }
                        if (lane == 0) g_d1 = dd;
                    }
                    __syncthreads();
                    const int d1 = g_d1;
                    // ---- build U for the group: cells 0 .. wr-1 real, cells wr, wr+1 zero ----
                    {
                        const int ncell = max(g_wr[0], g_wr[1]) + 2, per_side = 16 * ncell * 8;
#pragma unroll 4
                        for (int it = tid; it < 2 * per_side; it += kClThreads) {
                            const int side = it >= per_side, i2 = it - side * per_side;
                            const int q4 = i2 & 7, r2 = i2 >> 3, cell = r2 % ncell, ph = r2 / ncell;
                            const int wr = g_wr[side], x = g_win[side] + cell;
                            const bool real = cell < wr && x <= W - 1;
                            const AxisTap t0 = ytab[2 * ph], t1 = ytab[2 * ph + 1];
                            const float *fx = (side ? p.nhwcR : p.nhwcL) + img + (size_t)min(x, W - 1) * kClC + 4 * q4;
                            const float4 a0 = __ldg(reinterpret_cast<const float4 *>(fx + t0.olo));
                            const float4 a1 = __ldg(reinterpret_cast<const float4 *>(fx + t0.ohi));
                            float4 b0 = a0, b1 = a1;
                            if (t1.olo != t0.olo || t1.ohi != t0.ohi) {
                                b0 = (t1.olo == t0.ohi) ? a1 : __ldg(reinterpret_cast<const float4 *>(fx + t1.olo));
                                b1 = __ldg(reinterpret_cast<const float4 *>(fx + t1.ohi));
                            }
                            float4 u;
                            u.x = fmaf(t0.l, a1.x, t0.h * a0.x) + fmaf(t1.l, b1.x, t1.h * b0.x);
                            u.y = fmaf(t0.l, a1.y, t0.h * a0.y) + fmaf(t1.l, b1.y, t1.h * b0.y);
                            u.z = fmaf(t0.l, a1.z, t0.h * a0.z) + fmaf(t1.l, b1.z, t1.h * b0.z);
                            u.w = fmaf(t0.l, a1.w, t0.h * a0.w) + fmaf(t1.l, b1.w, t1.h * b0.w);
                            if (!real) u = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (cell <= wr + 1)
                                *reinterpret_cast<float4 *>(U + side * kClSideF + ph * kClRowF + cell * kClC + 4 * q4) = u;
                        }
                    }
                    for (int ds = d0; ds <= d1; ds += kClSub) {            // sub-batches whose x-sample tables fit shared memory
                        const int de = min(ds + kClSub - 1, d1);
                        // x-sample tables of slices ds .. de: thread = (slice, view, bin column)
                        if (tid < (de - ds + 1) * 32) {
                            const int sl = tid >> 5, side = (tid >> 4) & 1, pw = tid & 15;
                            const float4 g4 = geo[ds + sl];
                            const int wr = g_wr[side], win = g_win[side];
                            float4 e = make_float4(__uint_as_float((uint32_t)(wr * 128)), 0.f, __uint_as_float((uint32_t)(wr * 128)), 0.f);
                            if (pw >= t * PWT && pw < (t + 1) * PWT) {
                                const AxisSample s0 = axis_sample(side ? g4.z : g4.x, side ? g4.w : g4.y, pw, 0, W);
                                const AxisSample s1 = axis_sample(side ? g4.z : g4.x, side ? g4.w : g4.y, pw, 1, W);
                                if (s0.lo >= 0) { e.x = __uint_as_float((uint32_t)((s0.lo - win) * 128)); e.y = 0.25f * s0.l; }
                                if (s1.lo >= 0) { e.z = __uint_as_float((uint32_t)((s1.lo - win) * 128)); e.w = 0.25f * s1.l; }
                            }
                            xtab[sl][side][pw] = e;
                        }
                        __syncthreads();                                   // publishes the tables (and, first time round, U)
                        for (int d = ds; d <= de; ++d) {
                            const float4 (*xt)[16] = xtab[d - ds];
                            const size_t vox_row = ((size_t)d * 16 + warp) * 16;      // first voxel of this warp's bin row in slice d
                            if (phase == 0) {
                                float2 L[8], R[8];
                                float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                                for (int st = 0; st < 8; ++st) {
                                    const int pw = 2 * st + hw;
                                    L[st] = cl_bin(urow, xt[0][pw]);
                                    R[st] = cl_bin(urow + kClSideF * 4, xt[1][pw]);
                                    s0 = fmaf(L[st].x, L[st].x, fmaf(L[st].y, L[st].y, s0));
                                    s1 = fmaf(R[st].x, R[st].x, fmaf(R[st].y, R[st].y, s1));
                                    s2 = fmaf(L[st].x, R[st].x, fmaf(L[st].y, R[st].y, s2));
                                    amax = fmaxf(fmaxf(amax, fmaxf(fabsf(L[st].x), fabsf(L[st].y))), fmaxf(fabsf(R[st].x), fabsf(R[st].y)));
                                }
                                s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
                                float (*pt)[4] = part[d & 1];
                                if (lane == 0) { pt[warp][0] = s0; pt[warp][1] = s1; pt[warp][2] = s2; }
                                __syncthreads();
                                const float xc = cl_gate(pt, lane);
                                if (tid == 0 && v.xcross) v.xcross[(size_t)n * D + d] = xc;
                                const float g = p.gate ? xc : 1.0f;
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    cl_emit_chunk<2, OC>(L + 2 * q, R + 2 * q, g, wstage + ((cc + q) & 1) * 2 * kClChunkB, ghi, glo,
                                                         vox_row + 4 * q, (uint32_t)(4 * OC * 2), lane, hw);
                                cc += 4;
                            } else {
                                // tiled: bin columns [t PWT, (t+1) PWT) of the row, two per step (one per step when PWT == 1: both
                                // half-warps then hold the same bin and only the first counts)
                                const int p0 = t * PWT, nst = (PWT + 1) / 2;
                                const float once = (PWT == 1 && hw) ? 0.f : 1.f;
                                float s0 = 0.f, s1 = 0.f, s2 = 0.f;
                                for (int st = 0; st < nst; st += 2) {        // chunks of <= 2 steps = 4 voxels
                                    float2 l[2], r[2];
#pragma unroll
                                    for (int j = 0; j < 2; ++j) {
                                        const int pw = min(p0 + 2 * (st + j) + (PWT == 1 ? 0 : hw), p0 + PWT - 1);
                                        l[j] = cl_bin(urow, xt[0][pw]);
                                        r[j] = cl_bin(urow + kClSideF * 4, xt[1][pw]);
                                        if (st + j < nst) {
                                            s0 = fmaf(l[j].x, l[j].x, fmaf(l[j].y, l[j].y, s0));
                                            s1 = fmaf(r[j].x, r[j].x, fmaf(r[j].y, r[j].y, s1));
                                            s2 = fmaf(l[j].x, r[j].x, fmaf(l[j].y, r[j].y, s2));
                                            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(l[j].x), fabsf(l[j].y))), fmaxf(fabsf(r[j].x), fabsf(r[j].y)));
                                        }
                                    }
                                    if (phase == 2) {
                                        const float g = p.gate ? gtab[d] : 1.0f;
                                        unsigned char *buf = wstage + (cc & 1) * 2 * kClChunkB;
                                        const int nv = min(4, PWT - 2 * st);                     // voxels of this chunk: 4, 2 or 1
                                        if (nv == 4) cl_emit_chunk<2, OC>(l, r, g, buf, ghi, glo, vox_row + p0 + 2 * st, (uint32_t)(4 * OC * 2), lane, hw);
                                        else cl_emit_chunk<1, OC>(l, r, g, buf, ghi, glo, vox_row + p0 + 2 * st, (uint32_t)(nv * OC * 2), lane, hw);
                                        ++cc;
                                    }
                                }
                                if (phase == 1) {
                                    s0 = warp_sum(s0 * once); s1 = warp_sum(s1 * once); s2 = warp_sum(s2 * once);
                                    if (lane == 0) { partD[d][warp][0] += s0; partD[d][warp][1] += s1; partD[d][warp][2] += s2; }
                                }
                            }
                        }
                        __syncthreads();               // everyone is done with the tables (and, after the last sub-batch, with U)
                    }
                    d0 = d1 + 1;
                }
            }
            if (phase == 1) {
                // gate scalar of every slice from the per-warp sums of all tiles; the staging area goes back to its own use
                __syncthreads();
                for (int d = warp; d < D; d += 16) {
                    const float xc = cl_gate(partD[d], lane);
                    if (lane == 0) {
                        gtab[d] = xc;
                        if (v.xcross) v.xcross[(size_t)n * D + d] = xc;
                    }
                }
                __syncthreads();
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
    // |L - R| can reach twice the tracked maximum: report it when that is what would saturate
    if (p.rs) range_commit(p.rs, (OC == 96 && 2.f * amax >= 65504.f) ? 2.f * amax : amax);
}

}  // namespace side

using namespace side;

extern "C" size_t side_inst_costvol_cl_ws_bytes(int B, int C, int H, int W)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return sizeof(float) * 2 * (size_t)B * C * H * W + 256;      // channels-last copies of both feature maps + the work counter
}

extern "C" int side_inst_costvol_fwd_cl(const float *featL, const float *featR, const float *left, const float *right,
                                        const float *fb, const uint8_t *valid, void *cost_hi, void *cost_lo, float *depth_bin,
                                        float *xcross, int N, int B, int C, int H, int W, int D, int P, float x_clamp, int flags,
                                        void *ws, size_t ws_bytes, void *stream)
{
    SIDE_REQUIRE(N >= 0 && B > 0 && H > 1 && W > 1, "side_inst_costvol_fwd_cl: bad shape");
    SIDE_REQUIRE(C == kClC && P == 16, "side_inst_costvol_fwd_cl: built for C == 32 channels per view and P == 16 (got C=%d, P=%d)", C, P);
    SIDE_REQUIRE(D >= 2 && D <= kClMaxD, "side_inst_costvol_fwd_cl: D must be in 2..%d", kClMaxD);
    SIDE_REQUIRE((long long)B * H * W * C < (1ll << 31) && W < 32768, "side_inst_costvol_fwd_cl: features too large");
    SIDE_REQUIRE((flags & ~(SIDE_VOL_GATE | SIDE_VOL_FEAT_NHWC | SIDE_VOL_NO_DIFF)) == 0,
                 "side_inst_costvol_fwd_cl: valid flags are SIDE_VOL_GATE, SIDE_VOL_FEAT_NHWC, SIDE_VOL_NO_DIFF");
    if (N == 0) return SIDE_OK;
    SIDE_REQUIRE_DEV(featL); SIDE_REQUIRE_DEV(featR); SIDE_REQUIRE_DEV(left); SIDE_REQUIRE_DEV(right); SIDE_REQUIRE_DEV(fb);
    SIDE_REQUIRE_DEV(cost_hi); SIDE_REQUIRE_DEV(cost_lo); SIDE_REQUIRE_DEV(depth_bin);
    if (xcross) SIDE_REQUIRE_DEV(xcross);
    if (valid) SIDE_REQUIRE_DEV(valid);
    if (ws == nullptr || ws_bytes < side_inst_costvol_cl_ws_bytes(B, C, H, W) || !is_device_ptr(ws) ||
        (reinterpret_cast<uintptr_t>(ws) & 15)) {
        set_error("side_inst_costvol_fwd_cl: needs side_inst_costvol_cl_ws_bytes(...) bytes of 16-byte aligned device workspace");
        return SIDE_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t fsz = (size_t)B * C * H * W;
    float *nl = reinterpret_cast<float *>(ws), *nr = nl + fsz;
    int *counter = reinterpret_cast<int *>(nr + fsz);
    int rc;
    if (flags & SIDE_VOL_FEAT_NHWC) {                       // the features arrive channels-last [B, H, W, C]: no staging copies
        nl = const_cast<float *>(featL); nr = const_cast<float *>(featR);
    } else {
        if ((rc = launch_nchw_to_nhwc(featL, nl, B, C, H * W, st))) return rc;
        if ((rc = launch_nchw_to_nhwc(featR, nr, B, C, H * W, st))) return rc;
    }
    SIDE_CUDA(cudaMemsetAsync(counter, 0, sizeof(int), st));
    ClParams p{};
    p.v = VolParams{featL, featR, left, right, fb, valid, nullptr, depth_bin, xcross, nullptr, nullptr, nullptr, N, B, C, H, W, D, P, x_clamp};
    p.nhwcL = nl; p.nhwcR = nr;
    p.hi = reinterpret_cast<__half *>(cost_hi); p.lo = reinterpret_cast<__half *>(cost_lo);
    p.counter = counter; p.gate = (flags & SIDE_VOL_GATE) ? 1 : 0; p.rs = range_slot_next();
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        SIDE_CUDA(cudaGetDevice(&dev));
        SIDE_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    if (flags & SIDE_VOL_NO_DIFF) {
        if ((rc = set_smem_attr((const void *)inst_costvol_cl_kernel<64>, kClSmem))) return rc;
        inst_costvol_cl_kernel<64><<<(unsigned)std::min(N, sm_count), kClThreads, kClSmem, st>>>(p);
    } else {
        if ((rc = set_smem_attr((const void *)inst_costvol_cl_kernel<96>, kClSmem))) return rc;
        inst_costvol_cl_kernel<96><<<(unsigned)std::min(N, sm_count), kClThreads, kClSmem, st>>>(p);
    }
    SIDE_LAUNCH_CHECK("inst_costvol_cl_kernel");
    return SIDE_OK;
}
