// decode.cu -- CenterNet heat-map decode in ONE launch.
//
// Replaces ~25 ATen launches of the reference: _nms (decode.py:9-15), _topk (:17-33) -- two torch.topk
// calls plus gathers -- and _transpose_and_gather_feat (utils.py:21-26), which makes a full NHWC copy of
// every head (incl. the 168-channel kept_type map) just to fetch K rows.
//
// Design: grid = B*Cat CTAs of 1024 threads.  CTA (b, c) turns its class plane into order-preserving
// uint32 keys in shared memory (sigmoid + 3x3 max-pool NMS fused into the key computation, neighbours
// come straight from L2), finds the K-th largest key with a 4-pass 8-bit radix select (warp-aggregated
// shared-memory histogram atomics), collects the winners, bitonic-sorts them and parks them in a small
// workspace.  The last CTA of each image to finish (threadfence + ticket) merges the Cat*K candidates,
// then gathers the regression heads at the K winning pixels directly from NCHW -- no transposed copies.
// Tie rule everywhere: larger score first, then lower flat index (documented; torch.topk leaves it open).
#include <algorithm>
#include "common.cuh"

namespace side {

constexpr int kDecThreads = 1024;

__device__ __forceinline__ uint32_t f2key(float v)
{
    if (v == 0.f) v = 0.f;  // -0 -> +0 so that they tie like torch's comparison does
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// descending bitonic sort of n (power of two) u64 in shared memory, whole block participates
__device__ void bitonic_desc(unsigned long long *a, int n)
{
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = a[i], y = a[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (x < y) : (x > y)) { a[i] = y; a[ixj] = x; }
                }
            }
        }
    __syncthreads();
}

struct DecParams {
    const float *heat, *wh, *reg, *kept, *dim, *orien;
    float *bbox, *bbox_right;  // bbox flavour
    uint8_t *keep;
    int32_t *slot, *count;
    float *det, *det_right, *info;  // ddd flavour
    float *score;
    int32_t *ind, *cls;
    unsigned long long *ws_comp;  // [B*Cat*K]
    unsigned int *ws_ticket;      // [B]
    int B, Cat, H, W, K, Kpad, Mpad, grid;
    float wh_scale;
    int heat_is_logit;
    size_t regionA_bytes;
};

template <bool DDD>
__global__ void __launch_bounds__(kDecThreads) nms_topk_decode_kernel(DecParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *keys = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned long long *merge = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned long long *comp = reinterpret_cast<unsigned long long *>(smem_raw + p.regionA_bytes);
    __shared__ unsigned int hist[256];
    __shared__ unsigned int whist[kDecThreads / 32][256];   // per-warp private histograms: no cross-warp atomic contention
    __shared__ unsigned int sh_prefix, sh_remaining, sh_ngt, sh_running, sh_last;
    __shared__ unsigned int warp_cnt[32];

    const int b = blockIdx.x / p.Cat, c = blockIdx.x % p.Cat;
    const int H = p.H, W = p.W, HW = H * W, K = p.K;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float *hp = p.heat + ((size_t)b * p.Cat + c) * HW;

    // ---- 1. keys = order-preserving bits of heat * (maxpool3x3(heat) == heat) ------------------------
    for (int q = tid; q < HW; q += blockDim.x) {
        const int y = q / W, x = q - y * W;
        const float v = __ldg(hp + q);
        float m = v;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                m = fmaxf(m, __ldg(hp + yy * W + xx));
            }
        }
        float hv = v, hm = m;
        if (p.heat_is_logit) {  // sigmoid is monotone: max(sigmoid(x_i)) == sigmoid(max x_i)
            hv = sigmoid_acc(v);
            hm = (m == v) ? hv : sigmoid_acc(m);
        }
        keys[q] = f2key(hm == hv ? hv : hv * 0.0f);
    }
    if (tid == 0) { sh_prefix = 0u; sh_remaining = (unsigned)K; sh_ngt = 0u; sh_running = 0u; }
    __syncthreads();

    // ---- 2. radix select: K-th largest key ------------------------------------------------------------
    uint32_t mask = 0u;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < (kDecThreads / 32) * 256; i += blockDim.x) (&whist[0][0])[i] = 0u;
        __syncthreads();
        const uint32_t prefix = sh_prefix;
        for (int base = 0; base < HW; base += blockDim.x) {
            const int q = base + tid;
            bool valid = false;
            uint32_t bin = 0;
            if (q < HW) {
                const uint32_t k = keys[q];
                valid = (k & mask) == prefix;
                bin = (k >> shift) & 255u;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, valid ? bin : (256u + lane));
            if (valid && lane == (__ffs(peers) - 1)) atomicAdd(&whist[wid][bin], (unsigned)__popc(peers));
        }
        __syncthreads();
        if (tid < 256) {
            unsigned h = 0;
#pragma unroll 8
            for (int w = 0; w < kDecThreads / 32; ++w) h += whist[w][tid];
            hist[tid] = h;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned rem = sh_remaining, cum = 0;
            int sel = 0;
            for (int bin = 255; bin >= 0; --bin) {
                const unsigned h = hist[bin];
                if (cum + h >= rem) { sel = bin; break; }
                cum += h;
            }
            sh_remaining = rem - cum;
            sh_prefix = prefix | ((uint32_t)sel << shift);
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    const uint32_t T = sh_prefix;
    const unsigned need_eq = sh_remaining;         // how many keys == T are taken (lowest indices first)
    const unsigned n_gt = (unsigned)K - need_eq;   // keys strictly above T

    // ---- 3. collect winners ---------------------------------------------------------------------------
    for (int base = 0; base < HW; base += blockDim.x) {
        const int q = base + tid;
        if (q < HW) {
            const uint32_t k = keys[q];
            if (k > T) {
                const unsigned s = atomicAdd(&sh_ngt, 1u);
                comp[s] = ((unsigned long long)k << 32) | (unsigned long long)(0xffffffffu - (uint32_t)q);
            }
        }
    }
    for (int base = 0; base < HW; base += blockDim.x) {
        __syncthreads();
        const unsigned running = sh_running;
        if (running >= need_eq) break;  // uniform
        const int q = base + tid;
        const bool f = q < HW && keys[q] == T;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            const unsigned cw = warp_cnt[w];
            if (w < wid) before += cw;
            total += cw;
        }
        const unsigned rank = running + before + __popc(bal & ((1u << lane) - 1u));
        if (f && rank < need_eq)
            comp[n_gt + rank] = ((unsigned long long)T << 32) | (unsigned long long)(0xffffffffu - (uint32_t)q);
        __syncthreads();
        if (tid == 0) sh_running = running + total;
    }
    __syncthreads();
    for (int i = K + tid; i < p.Kpad; i += blockDim.x) comp[i] = 0ull;
    bitonic_desc(comp, p.Kpad);

    // ---- 4. park per-class winners, elect the merging CTA ------------------------------------------------
    unsigned long long *wsc = p.ws_comp + ((size_t)b * p.Cat + c) * K;
    for (int i = tid; i < K; i += blockDim.x) wsc[i] = comp[i];
    __threadfence();
    __syncthreads();
    if (tid == 0) sh_last = (atomicAdd(&p.ws_ticket[b], 1u) == (unsigned)(p.Cat - 1)) ? 1u : 0u;
    __syncthreads();
    if (!sh_last) return;
    __threadfence();

    // ---- 5. merge Cat*K candidates (position = c*K + rank breaks ties, as in topk over the [Cat*K] view) --
    const int M = p.Cat * K;
    const unsigned long long *wsb = p.ws_comp + (size_t)b * M;
    for (int i = tid; i < p.Mpad; i += blockDim.x) {
        unsigned long long e = 0ull;
        if (i < M) e = (__ldcg(wsb + i) & 0xffffffff00000000ull) | (unsigned long long)(0xffffffffu - (uint32_t)i);
        merge[i] = e;
    }
    bitonic_desc(merge, p.Mpad);

    // ---- 6. gathers ---------------------------------------------------------------------------------------
    int my_keep = 0;
    if (tid < K) {
        const unsigned long long e = merge[tid];
        const uint32_t pos = 0xffffffffu - (uint32_t)(e & 0xffffffffull);
        const int cls = (int)(pos / (unsigned)K);
        const int idx = (int)(0xffffffffu - (uint32_t)(__ldcg(wsb + pos) & 0xffffffffull));
        const float sc = key2f((uint32_t)(e >> 32));
        const float xs = (float)(idx % W), ys = (float)(idx / W);
        const size_t bk = (size_t)b * K + tid;
        if (p.score) p.score[bk] = sc;
        if (p.ind) p.ind[bk] = idx;
        if (p.cls) p.cls[bk] = cls;
        const float r0 = __ldg(p.reg + ((size_t)b * 3 + 0) * HW + idx);
        const float r1 = __ldg(p.reg + ((size_t)b * 3 + 1) * HW + idx);
        const float r2 = __ldg(p.reg + ((size_t)b * 3 + 2) * HW + idx);
        float w0 = __ldg(p.wh + ((size_t)b * 3 + 0) * HW + idx);
        float w1 = __ldg(p.wh + ((size_t)b * 3 + 1) * HW + idx);
        float w2 = __ldg(p.wh + ((size_t)b * 3 + 2) * HW + idx);
        const float cx = __fadd_rn(xs, r0), cxr = __fadd_rn(xs, r1), cy = __fadd_rn(ys, r2);
        if (!DDD) {
            w0 = __fmul_rn(w0, p.wh_scale); w1 = __fmul_rn(w1, p.wh_scale); w2 = __fmul_rn(w2, p.wh_scale);
            const float h0 = __fmul_rn(0.5f, w0), h1 = __fmul_rn(0.5f, w1), h2 = __fmul_rn(0.5f, w2);
            float *o = p.bbox + bk * 5, *r = p.bbox_right + bk * 5;
            const float o1 = __fsub_rn(cx, h0), o2 = __fsub_rn(cy, h2), o3 = __fadd_rn(cx, h0), o4 = __fadd_rn(cy, h2);
            o[0] = (float)b; o[1] = o1; o[2] = o2; o[3] = o3; o[4] = o4;
            r[0] = (float)b; r[1] = __fsub_rn(cxr, h1); r[2] = o2; r[3] = __fadd_rn(cxr, h1); r[4] = o4;
            my_keep = __fadd_rn(__fadd_rn(__fadd_rn(o1, o2), o3), o4) > 0.f ? 1 : 0;
            if (p.keep) p.keep[bk] = (uint8_t)my_keep;
        } else {
            float *d = p.det + bk * 6, *dr = p.det_right + bk * 6, *f = p.info + bk * 9;
            d[0] = cx; d[1] = cy; d[2] = w0; d[3] = w2; d[4] = sc; d[5] = (float)cls;
            dr[0] = cxr; dr[1] = cy; dr[2] = w1; dr[3] = w2; dr[4] = sc; dr[5] = (float)cls;
            f[0] = __ldg(p.dim + ((size_t)b * 3 + 0) * HW + idx);
            f[1] = __ldg(p.dim + ((size_t)b * 3 + 1) * HW + idx);
            f[2] = __ldg(p.dim + ((size_t)b * 3 + 2) * HW + idx);
            f[3] = __ldg(p.orien + ((size_t)b * 2 + 0) * HW + idx);
            f[4] = __ldg(p.orien + ((size_t)b * 2 + 1) * HW + idx);
        }
    }
    if (!DDD) {
        // slot = exclusive rank among this image's kept rows, count = total kept (K <= blockDim.x)
        const unsigned bal = __ballot_sync(0xffffffffu, my_keep);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            const unsigned cw = warp_cnt[w];
            if (w < wid) before += cw;
            total += cw;
        }
        if (tid < K && p.slot) p.slot[(size_t)b * K + tid] = (int)(before + __popc(bal & ((1u << lane) - 1u)));
        if (tid == 0 && p.count) p.count[b] = (int)total;
    } else {
        // kept_type argmaxes: one warp per detection, lanes stride the channel range; first max wins
        const int g = p.grid;
        for (int k = wid; k < K; k += (int)(blockDim.x >> 5)) {
            const unsigned long long e = merge[k];
            const uint32_t pos = 0xffffffffu - (uint32_t)(e & 0xffffffffull);
            const int idx = (int)(0xffffffffu - (uint32_t)(__ldcg(wsb + pos) & 0xffffffffull));
            const float *kp = p.kept + (size_t)b * 6 * g * HW + idx;
            int res[3];
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int c0 = part == 0 ? 0 : (part == 1 ? 4 * g : 5 * g);
                const int nc = part == 0 ? 4 * g : g;
                float bv = -INFINITY;
                int bi = 0x7fffffff;
                for (int j = lane; j < nc; j += 32) {
                    const float v = __ldg(kp + (size_t)(c0 + j) * HW);
                    if (bi == 0x7fffffff || v > bv) { bv = v; bi = j; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                res[part] = bi;
            }
            if (lane == 0) {
                float *f = p.info + ((size_t)b * K + k) * 9;
                f[5] = (float)res[1];
                f[6] = (float)res[2];
                f[7] = (float)(res[0] % g);
                f[8] = (float)(res[0] / g);
            }
        }
    }
}

static int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

static int launch_decode(DecParams &p, bool ddd, void *ws, size_t ws_bytes, void *stream)
{
    SIDE_REQUIRE(p.B > 0 && p.Cat > 0 && p.H > 0 && p.W > 0, "decode: bad shape");
    SIDE_REQUIRE(p.K >= 1 && p.K <= kDecThreads && p.K <= p.H * p.W, "decode: K=%d out of range (1..min(%d, H*W))", p.K,
                 kDecThreads);
    SIDE_REQUIRE((long long)p.Cat * p.K <= 8192, "decode: Cat*K=%lld exceeds 8192", (long long)p.Cat * p.K);
    const size_t need = side_decode_ws_bytes(p.B, p.Cat, p.K);
    if (ws == nullptr || ws_bytes < need) {
        set_error("decode: workspace too small (%zu < %zu)", ws_bytes, need);
        return SIDE_ERR_WORKSPACE;
    }
    SIDE_REQUIRE_DEV(ws);
    p.Kpad = next_pow2(p.K);
    p.Mpad = next_pow2(p.Cat * p.K);
    size_t regionA = (size_t)p.H * p.W * 4;
    if ((size_t)p.Mpad * 8 > regionA) regionA = (size_t)p.Mpad * 8;
    regionA = (regionA + 15) & ~(size_t)15;
    p.regionA_bytes = regionA;
    const size_t smem = regionA + (size_t)p.Kpad * 8;
    {   // static (per-warp histograms, ~33 KB) + dynamic shared memory against the 227 KB opt-in limit of sm_100a
        cudaFuncAttributes fa;
        SIDE_CUDA(cudaFuncGetAttributes(&fa, ddd ? (const void *)nms_topk_decode_kernel<true> : (const void *)nms_topk_decode_kernel<false>));
        if (smem + fa.sharedSizeBytes > 227 * 1024) {
            set_error("decode: H*W=%d needs %zu + %zu bytes of shared memory (> 227 KB)", p.H * p.W, smem, (size_t)fa.sharedSizeBytes);
            return SIDE_ERR_UNSUPPORTED;
        }
    }
    p.ws_ticket = reinterpret_cast<unsigned int *>(ws);
    p.ws_comp = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(ws) + 256 * ((p.B * 4 + 255) / 256));
    cudaStream_t st = (cudaStream_t)stream;
    SIDE_CUDA(cudaMemsetAsync(p.ws_ticket, 0, sizeof(unsigned int) * p.B, st));
    int rc;
    // the kernel also has ~33 KB of STATIC shared memory (per-warp histograms): static + dynamic can exceed the 48 KB
    // default even when the dynamic part alone does not, so always opt in
    const size_t attr = std::max(smem, (size_t)64 * 1024);
    if (ddd) {
        if ((rc = set_smem_attr((const void *)nms_topk_decode_kernel<true>, attr))) return rc;
        nms_topk_decode_kernel<true><<<p.B * p.Cat, kDecThreads, smem, st>>>(p);
    } else {
        if ((rc = set_smem_attr((const void *)nms_topk_decode_kernel<false>, attr))) return rc;
        nms_topk_decode_kernel<false><<<p.B * p.Cat, kDecThreads, smem, st>>>(p);
    }
    SIDE_LAUNCH_CHECK("nms_topk_decode_kernel");
    return SIDE_OK;
}

}  // namespace side

using namespace side;

extern "C" size_t side_decode_ws_bytes(int B, int Cat, int K)
{
    if (B <= 0 || Cat <= 0 || K <= 0) return 0;
    return (size_t)256 * ((B * 4 + 255) / 256) + sizeof(unsigned long long) * (size_t)B * Cat * K;
}

extern "C" int side_bbox_decode(const float *heat, const float *wh, const float *reg, float *bbox, float *bbox_right,
                                uint8_t *keep, int32_t *slot, int32_t *count, float *score, int32_t *ind,
                                int32_t *cls, int B, int Cat, int H, int W, int K, float wh_scale, int flags, void *ws,
                                size_t ws_bytes, void *stream)
{
    SIDE_REQUIRE_DEV(heat); SIDE_REQUIRE_DEV(wh); SIDE_REQUIRE_DEV(reg); SIDE_REQUIRE_DEV(bbox); SIDE_REQUIRE_DEV(bbox_right);
    DecParams p{};
    p.heat = heat; p.wh = wh; p.reg = reg; p.bbox = bbox; p.bbox_right = bbox_right; p.keep = keep; p.slot = slot;
    p.count = count; p.score = score; p.ind = ind; p.cls = cls;
    p.B = B; p.Cat = Cat; p.H = H; p.W = W; p.K = K; p.wh_scale = wh_scale; p.grid = 1;
    p.heat_is_logit = (flags & SIDE_DECODE_HEAT_IS_LOGIT) ? 1 : 0;
    return launch_decode(p, false, ws, ws_bytes, stream);
}

extern "C" int side_ddd_decode(const float *heat, const float *kept, const float *dim, const float *orien,
                               const float *wh, const float *reg, float *det, float *det_right, float *info,
                               float *score, int32_t *ind, int32_t *cls, int B, int Cat, int H, int W, int grid, int K,
                               int flags, void *ws, size_t ws_bytes, void *stream)
{
    SIDE_REQUIRE_DEV(heat); SIDE_REQUIRE_DEV(kept); SIDE_REQUIRE_DEV(dim); SIDE_REQUIRE_DEV(orien); SIDE_REQUIRE_DEV(wh);
    SIDE_REQUIRE_DEV(reg); SIDE_REQUIRE_DEV(det); SIDE_REQUIRE_DEV(det_right); SIDE_REQUIRE_DEV(info);
    SIDE_REQUIRE(grid >= 1, "side_ddd_decode: grid must be >= 1");
    DecParams p{};
    p.heat = heat; p.kept = kept; p.dim = dim; p.orien = orien; p.wh = wh; p.reg = reg;
    p.det = det; p.det_right = det_right; p.info = info; p.score = score; p.ind = ind; p.cls = cls;
    p.B = B; p.Cat = Cat; p.H = H; p.W = W; p.K = K; p.wh_scale = 1.0f; p.grid = grid;
    p.heat_is_logit = (flags & SIDE_DECODE_HEAT_IS_LOGIT) ? 1 : 0;
    return launch_decode(p, true, ws, ws_bytes, stream);
}
