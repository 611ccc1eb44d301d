// voxel.cu -- instance voxel volume of the stereo_network_new variant.
//
// Reference: src/lib/models/networks/stereo_network_new.py.  get_voxel (:160-283) is a host-side Python double loop
// (images x RoIs) that builds, per RoI, a 10 x 10 x 10 metric grid (0.5 m x 0.5 m x 1 m cells) around the 3-D point the
// box pair triangulates to, projects it into both views (project_rect_to_image :34-44, then the `trans` affine to feature
// pixels) and normalises it for F.grid_sample; forward (:409-449) then samples the 64-channel reduced features of both
// views per image (two grid_sample calls + masking + three tensor copies) into voxel = cat(L - R, L, R) [N, 192, 1000].
// Here:
//   voxel_coords_kernel  get_voxel's seven outputs, one thread per voxel (drop-in for the function; no host loop).
//   voxel_volume_kernel  the fused forward path: CTA = (RoI, 128 voxels).  Sample geometry once per voxel and view; the
//                        gather reads the channels-last feature copy with lanes over channel pairs (a corner = 256
//                        contiguous bytes per warp), values are transposed through shared memory and every plane of
//                        the output (L - R, L, R per channel) leaves as 512 contiguous bytes.  grid_sample semantics:
//                        bilinear, padding_mode='zeros', align_corners selectable (default False = what the reference
//                        computes under the torch in this image); invalid voxels sample nothing and write zeros.
//   voxel_volume_bwd_kernel  adjoint (training): gL = g[L] + g[L-R], gR = g[R] - g[L-R], scattered with the forward
//                        weights into channels-last gradient buffers (fp32 atomics), transposed back by the caller.
#include <algorithm>

#include "common.cuh"

namespace side {

constexpr int kVxRes = 10, kVxN = 1000;       // voxels per RoI
constexpr int kVxTile = 128, kVxThreads = 256;

struct VoxelParams {
    const float *left, *right;       // [N, 5] boxes (b, x1, y1, x2, y2) in feature pixels
    const float *p2, *p3;            // [B, 3, 4]
    const float *fb;                 // [B]
    const float *trans, *trans_inv;  // [B, 2, 3]
    int N, B, C, H, W;
    float u_max, v_max;              // input_w / 4 - 1, input_h / 4 - 1 (stereo_network_new.py:243-244)
    int align;
};

struct VoxelGeom {
    float uf, vf, z;                 // left-view feature coordinates + metric depth of the voxel
    float ufr, vfr;                  // right view
};

// float32 restatement of get_voxel's per-RoI set-up (:176-213); every operation one rounded float op, sums left to right
struct VoxelRoi {
    float x, y, z, depth;
    int b;
};

__device__ inline VoxelRoi voxel_roi(const VoxelParams &p, int n)
{
    const float *lb = p.left + 5 * n, *rb = p.right + 5 * n;
    VoxelRoi r;
    r.b = min(max((int)lb[0], 0), p.B - 1);
    const float *ti = p.trans_inv + 6 * r.b, *P2 = p.p2 + 12 * r.b;
    // pt = [x, y, 1] @ trans_inv^T
    auto tx = [&](float x, float y) { return __fadd_rn(__fadd_rn(__fmul_rn(x, ti[0]), __fmul_rn(y, ti[1])), ti[2]); };
    auto ty = [&](float x, float y) { return __fadd_rn(__fadd_rn(__fmul_rn(x, ti[3]), __fmul_rn(y, ti[4])), ti[5]); };
    const float cx = __fdiv_rn(__fadd_rn(tx(lb[1], lb[2]), tx(lb[3], lb[4])), 2.f);
    const float cy = __fdiv_rn(__fadd_rn(ty(lb[1], lb[2]), ty(lb[3], lb[4])), 2.f);
    const float cxr = __fdiv_rn(__fadd_rn(tx(rb[1], rb[2]), tx(rb[3], rb[4])), 2.f);
    r.depth = __fdiv_rn(p.fb[r.b], __fsub_rn(cx, cxr));
    r.z = __fsub_rn(r.depth, P2[11]);
    r.x = __fdiv_rn(__fsub_rn(__fsub_rn(__fmul_rn(cx, r.depth), P2[3]), __fmul_rn(P2[2], r.z)), P2[0]);
    r.y = __fdiv_rn(__fsub_rn(__fsub_rn(__fmul_rn(cy, r.depth), P2[7]), __fmul_rn(P2[6], r.z)), P2[5]);
    return r;
}

// voxel v = (ix, iy, iz) of torch.meshgrid(xs, ys, zs) (indexing 'ij'): projected into both views
__device__ inline VoxelGeom voxel_geom(const VoxelParams &p, const VoxelRoi &r, int v)
{
    const int ix = v / 100, iy = (v / 10) % 10, iz = v % 10;
    const float X = __fadd_rn(__fadd_rn(-2.5f + 0.5f * (float)ix, 0.25f), r.x);
    const float Y = __fadd_rn(__fadd_rn(-2.5f + 0.5f * (float)iy, 0.25f), r.y);
    const float Z = __fadd_rn(__fadd_rn(-5.f + (float)iz, 0.5f), r.z);
    const float *tr = p.trans + 6 * r.b;
    VoxelGeom g;
    g.z = Z;
    auto project = [&](const float *P, float &uf, float &vf) {
        // [X, Y, Z, 1] @ P^T, then / w, then [u, v, 1] @ trans^T
        const float a = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, P[0]), __fmul_rn(Y, P[1])), __fmul_rn(Z, P[2])), P[3]);
        const float b = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, P[4]), __fmul_rn(Y, P[5])), __fmul_rn(Z, P[6])), P[7]);
        const float w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(X, P[8]), __fmul_rn(Y, P[9])), __fmul_rn(Z, P[10])), P[11]);
        const float u = __fdiv_rn(a, w), vv = __fdiv_rn(b, w), one = __fdiv_rn(w, w);
        uf = __fadd_rn(__fadd_rn(__fmul_rn(u, tr[0]), __fmul_rn(vv, tr[1])), __fmul_rn(one, tr[2]));
        vf = __fadd_rn(__fadd_rn(__fmul_rn(u, tr[3]), __fmul_rn(vv, tr[4])), __fmul_rn(one, tr[5]));
    };
    project(p.p2 + 12 * r.b, g.uf, g.vf);
    project(p.p3 + 12 * r.b, g.ufr, g.vfr);
    return g;
}

// (coord - 0) / (max - 0) * 2 - 1 and the [-1, 1] test (:246-263, :271-278)
__device__ __forceinline__ float vx_norm(float c, float cmax) { return __fsub_rn(__fmul_rn(__fdiv_rn(c, cmax), 2.f), 1.f); }

__global__ void voxel_coords_kernel(VoxelParams p, const float *__restrict__ depth_bins, int D, float *__restrict__ norm3,
                                    float *__restrict__ valid3, float *__restrict__ normL, float *__restrict__ validL,
                                    float *__restrict__ normR, float *__restrict__ validR, float *__restrict__ depth_ori)
{
    const int n = blockIdx.x;
    const VoxelRoi r = voxel_roi(p, n);
    if (threadIdx.x == 0) depth_ori[n] = r.depth;
    float dmin = 0.f, dmax = 0.f;
    if (depth_bins) {
        dmin = dmax = depth_bins[(size_t)n * D];
        for (int i = 1; i < D; ++i) {
            const float d = depth_bins[(size_t)n * D + i];
            dmin = fminf(dmin, d); dmax = fmaxf(dmax, d);
        }
    }
    const float *lb = p.left + 5 * n;
    for (int v = threadIdx.x; v < kVxN; v += blockDim.x) {
        const VoxelGeom g = voxel_geom(p, r, v);
        const size_t o = (size_t)n * kVxN + v;
        if (norm3) {
            const float a = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(g.uf, lb[1]), __fsub_rn(lb[3], lb[1])), 2.f), 1.f);
            const float b = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(g.vf, lb[2]), __fsub_rn(lb[4], lb[2])), 2.f), 1.f);
            const float c = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(g.z, dmin), __fsub_rn(dmax, dmin)), 2.f), 1.f);
            norm3[3 * o] = a; norm3[3 * o + 1] = b; norm3[3 * o + 2] = c;
            valid3[o] = (a >= -1.f && a <= 1.f && b >= -1.f && b <= 1.f && c >= -1.f && c <= 1.f) ? 1.f : 0.f;
        }
        const float lu = vx_norm(g.uf, p.u_max), lv = vx_norm(g.vf, p.v_max);
        const float ru = vx_norm(g.ufr, p.u_max), rv = vx_norm(g.vfr, p.v_max);
        normL[2 * o] = lu; normL[2 * o + 1] = lv;
        normR[2 * o] = ru; normR[2 * o + 1] = rv;
        validL[o] = (lu >= -1.f && lu <= 1.f && lv >= -1.f && lv <= 1.f) ? 1.f : 0.f;
        validR[o] = (ru >= -1.f && ru <= 1.f && rv >= -1.f && rv <= 1.f) ? 1.f : 0.f;
    }
}

// one view's bilinear taps of a voxel: element offsets of the four corners in the channels-last image (or -1) and weights
struct VxTap {
    int o[4];
    float w[4];
};

__device__ inline VxTap vx_tap(float gx, float gy, bool valid, int H, int W, int C, int align)
{
    VxTap t;
#pragma unroll
    for (int k = 0; k < 4; ++k) { t.o[k] = -1; t.w[k] = 0.f; }
    if (!valid) return t;                     // forward multiplies the samples of an invalid voxel by 0 (:440, :444)
    // grid_sampler_unnormalize (GridSampler.h), padding_mode = zeros
    const float x = align ? __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(W - 1))
                          : __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 2.f);
    const float y = align ? __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(H - 1))
                          : __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 2.f);
    const float xw = floorf(x), yn = floorf(y);
    const float we = __fsub_rn(x, xw), ww = __fsub_rn(1.f, we), ws_ = __fsub_rn(y, yn), wn = __fsub_rn(1.f, ws_);
    const int ix = (int)xw, iy = (int)yn;
    const float w4[4] = {__fmul_rn(wn, ww), __fmul_rn(wn, we), __fmul_rn(ws_, ww), __fmul_rn(ws_, we)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = ix + (k & 1), yy = iy + (k >> 1);
        if (xx >= 0 && xx < W && yy >= 0 && yy < H) { t.o[k] = (yy * W + xx) * C; t.w[k] = w4[k]; }
    }
    return t;
}

__device__ inline void vx_taps_of(const VoxelParams &p, const VoxelRoi &r, int v, VxTap &tl, VxTap &tr)
{
    const VoxelGeom g = voxel_geom(p, r, v);
    const float lu = vx_norm(g.uf, p.u_max), lv = vx_norm(g.vf, p.v_max);
    const float ru = vx_norm(g.ufr, p.u_max), rv = vx_norm(g.vfr, p.v_max);
    tl = vx_tap(lu, lv, lu >= -1.f && lu <= 1.f && lv >= -1.f && lv <= 1.f, p.H, p.W, p.C, p.align);
    tr = vx_tap(ru, rv, ru >= -1.f && ru <= 1.f && rv >= -1.f && rv <= 1.f, p.H, p.W, p.C, p.align);
}

// C == 64 (feaRuduce of the variant, stereo_network_new.py:319-323): lane = channel pair
__global__ void __launch_bounds__(kVxThreads) voxel_volume_kernel(VoxelParams p, const float *__restrict__ nhwcL,
                                                                 const float *__restrict__ nhwcR, float *__restrict__ voxel,
                                                                 float *__restrict__ depth_ori)
{
    extern __shared__ float sm[];                        // sL[64][129], sR[64][129]
    __shared__ VxTap tapL[kVxTile], tapR[kVxTile];
    const int n = blockIdx.y, v0 = blockIdx.x * kVxTile, nv = min(kVxTile, kVxN - v0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int C = 64, LD = kVxTile + 1;
    float *sL = sm, *sR = sm + C * LD;
    const VoxelRoi r = voxel_roi(p, n);
    if (blockIdx.x == 0 && tid == 0 && depth_ori) depth_ori[n] = r.depth;
    if (tid < nv) vx_taps_of(p, r, v0 + tid, tapL[tid], tapR[tid]);
    __syncthreads();
    const size_t img = (size_t)r.b * p.H * p.W * C;
    const float2 *fl = reinterpret_cast<const float2 *>(nhwcL + img) + lane, *fr = reinterpret_cast<const float2 *>(nhwcR + img) + lane;
    for (int v = warp; v < nv; v += kVxThreads / 32) {
        float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) {          // corner order nw, ne, sw, se (GridSamplerKernel.cpp)
            const int ol = tapL[v].o[k], orr = tapR[v].o[k];
            if (ol >= 0) {
                const float2 q = __ldg(fl + (ol >> 1));
                const float w = tapL[v].w[k];
                a.x = __fadd_rn(a.x, __fmul_rn(q.x, w)); a.y = __fadd_rn(a.y, __fmul_rn(q.y, w));
            }
            if (orr >= 0) {
                const float2 q = __ldg(fr + (orr >> 1));
                const float w = tapR[v].w[k];
                b.x = __fadd_rn(b.x, __fmul_rn(q.x, w)); b.y = __fadd_rn(b.y, __fmul_rn(q.y, w));
            }
        }
        sL[(2 * lane) * LD + v] = a.x; sL[(2 * lane + 1) * LD + v] = a.y;
        sR[(2 * lane) * LD + v] = b.x; sR[(2 * lane + 1) * LD + v] = b.y;
    }
    __syncthreads();
    // planes: [0, C) = L - R, [C, 2C) = L, [2C, 3C) = R (:447); each (plane, tile) = nv contiguous floats
    float *out = voxel + (size_t)n * 3 * C * kVxN + v0;
    for (int i = tid; i < C * kVxTile; i += kVxThreads) {
        const int c = i / kVxTile, v = i - c * kVxTile;
        if (v >= nv) continue;
        const float l = sL[c * LD + v], rr = sR[c * LD + v];
        st_cs(out + (size_t)c * kVxN + v, __fsub_rn(l, rr));
        st_cs(out + (size_t)(C + c) * kVxN + v, l);
        st_cs(out + (size_t)(2 * C + c) * kVxN + v, rr);
    }
}

__global__ void __launch_bounds__(kVxThreads) voxel_volume_bwd_kernel(VoxelParams p, const float *__restrict__ gvoxel,
                                                                     float *__restrict__ gnhwcL, float *__restrict__ gnhwcR)
{
    extern __shared__ float sm[];                        // gL[64][129], gR[64][129]
    __shared__ VxTap tapL[kVxTile], tapR[kVxTile];
    const int n = blockIdx.y, v0 = blockIdx.x * kVxTile, nv = min(kVxTile, kVxN - v0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int C = 64, LD = kVxTile + 1;
    float *sL = sm, *sR = sm + C * LD;
    const VoxelRoi r = voxel_roi(p, n);
    if (tid < nv) vx_taps_of(p, r, v0 + tid, tapL[tid], tapR[tid]);
    const float *g = gvoxel + (size_t)n * 3 * C * kVxN + v0;
    for (int i = tid; i < C * kVxTile; i += kVxThreads) {
        const int c = i / kVxTile, v = i - c * kVxTile;
        if (v >= nv) continue;
        const float gd = __ldg(g + (size_t)c * kVxN + v);
        sL[c * LD + v] = __ldg(g + (size_t)(C + c) * kVxN + v) + gd;
        sR[c * LD + v] = __ldg(g + (size_t)(2 * C + c) * kVxN + v) - gd;
    }
    __syncthreads();
    const size_t img = (size_t)r.b * p.H * p.W * C;
    for (int v = warp; v < nv; v += kVxThreads / 32) {
        const float a0 = sL[(2 * lane) * LD + v], a1 = sL[(2 * lane + 1) * LD + v];
        const float b0 = sR[(2 * lane) * LD + v], b1 = sR[(2 * lane + 1) * LD + v];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ol = tapL[v].o[k], orr = tapR[v].o[k];
            if (ol >= 0) {
                const float w = tapL[v].w[k];
                atomicAdd(gnhwcL + img + ol + 2 * lane, a0 * w);
                atomicAdd(gnhwcL + img + ol + 2 * lane + 1, a1 * w);
            }
            if (orr >= 0) {
                const float w = tapR[v].w[k];
                atomicAdd(gnhwcR + img + orr + 2 * lane, b0 * w);
                atomicAdd(gnhwcR + img + orr + 2 * lane + 1, b1 * w);
            }
        }
    }
}

static int vx_check(const VoxelParams &p)
{
    SIDE_REQUIRE(p.N >= 0 && p.B > 0 && p.H > 1 && p.W > 1, "side_voxel: bad shape");
    SIDE_REQUIRE((long long)p.B * p.H * p.W * p.C < (1ll << 31), "side_voxel: features too large");
    SIDE_REQUIRE(p.N <= 65535, "side_voxel: at most 65535 RoIs per call");
    return SIDE_OK;
}

}  // namespace side

using namespace side;

#define VX_PARAMS()                                                                                                        \
    VoxelParams p{left, right, p2, p3, fb, trans, trans_inv, N, B, C, H, W, (float)(input_w / 4.0 - 1.0),                  \
                  (float)(input_h / 4.0 - 1.0), (flags & SIDE_VOXEL_ALIGN_CORNERS) ? 1 : 0};                               \
    {                                                                                                                      \
        int rc0 = vx_check(p);                                                                                             \
        if (rc0) return rc0;                                                                                               \
    }                                                                                                                      \
    if (N == 0) return SIDE_OK;                                                                                            \
    SIDE_REQUIRE_DEV(left); SIDE_REQUIRE_DEV(right); SIDE_REQUIRE_DEV(p2); SIDE_REQUIRE_DEV(p3); SIDE_REQUIRE_DEV(fb);       \
    SIDE_REQUIRE_DEV(trans); SIDE_REQUIRE_DEV(trans_inv)

extern "C" int side_voxel_coords(const float *left, const float *right, const float *p2, const float *p3, const float *fb,
                                 const float *trans, const float *trans_inv, const float *depth_bins, int N, int B, int D, int H,
                                 int W, int input_h, int input_w, float *norm3, float *valid3, float *normL, float *validL,
                                 float *normR, float *validR, float *depth_ori, void *stream)
{
    const int C = 1, flags = 0;
    H = std::max(H, 2); W = std::max(W, 2);      // the feature size plays no role in the coordinates
    VX_PARAMS();
    SIDE_REQUIRE_DEV(normL); SIDE_REQUIRE_DEV(validL); SIDE_REQUIRE_DEV(normR); SIDE_REQUIRE_DEV(validR); SIDE_REQUIRE_DEV(depth_ori);
    if (norm3) {
        SIDE_REQUIRE_DEV(valid3);
        SIDE_REQUIRE_DEV(depth_bins);
        SIDE_REQUIRE(D > 0, "side_voxel_coords: D must be positive with depth_bins");
    }
    voxel_coords_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(p, norm3 ? depth_bins : nullptr, D, norm3, valid3, normL, validL, normR,
                                                            validR, depth_ori);
    SIDE_LAUNCH_CHECK("voxel_coords_kernel");
    return SIDE_OK;
}

extern "C" size_t side_voxel_volume_ws_bytes(int B, int C, int H, int W)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return sizeof(float) * 2 * (size_t)B * C * H * W;
}

extern "C" int side_voxel_volume_fwd(const float *featL, const float *featR, const float *left, const float *right, const float *p2,
                                     const float *p3, const float *fb, const float *trans, const float *trans_inv, float *voxel,
                                     float *depth_ori, int N, int B, int C, int H, int W, int input_h, int input_w, int flags,
                                     void *ws, size_t ws_bytes, void *stream)
{
    VX_PARAMS();
    SIDE_REQUIRE(C == 64, "side_voxel_volume_fwd: built for the variant's 64 reduced channels (got %d)", C);
    SIDE_REQUIRE_DEV(featL); SIDE_REQUIRE_DEV(featR); SIDE_REQUIRE_DEV(voxel);
    if (depth_ori) SIDE_REQUIRE_DEV(depth_ori);
    if (ws == nullptr || ws_bytes < side_voxel_volume_ws_bytes(B, C, H, W) || !is_device_ptr(ws)) {
        set_error("side_voxel_volume_fwd: needs side_voxel_volume_ws_bytes(...) bytes of device workspace");
        return SIDE_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float *nl = reinterpret_cast<float *>(ws), *nr = nl + (size_t)B * C * H * W;
    int rc;
    if ((rc = launch_nchw_to_nhwc(featL, nl, B, C, H * W, st))) return rc;
    if ((rc = launch_nchw_to_nhwc(featR, nr, B, C, H * W, st))) return rc;
    const size_t smem = sizeof(float) * 2 * 64 * (kVxTile + 1);
    if ((rc = set_smem_attr((const void *)voxel_volume_kernel, smem))) return rc;
    voxel_volume_kernel<<<dim3((kVxN + kVxTile - 1) / kVxTile, N), kVxThreads, smem, st>>>(p, nl, nr, voxel, depth_ori);
    SIDE_LAUNCH_CHECK("voxel_volume_kernel");
    return SIDE_OK;
}

extern "C" int side_voxel_volume_bwd(const float *gvoxel, const float *left, const float *right, const float *p2, const float *p3,
                                     const float *fb, const float *trans, const float *trans_inv, float *gfeatL, float *gfeatR,
                                     int N, int B, int C, int H, int W, int input_h, int input_w, int flags, void *ws,
                                     size_t ws_bytes, void *stream)
{
    VX_PARAMS();
    SIDE_REQUIRE(C == 64, "side_voxel_volume_bwd: built for the variant's 64 reduced channels (got %d)", C);
    SIDE_REQUIRE_DEV(gvoxel); SIDE_REQUIRE_DEV(gfeatL); SIDE_REQUIRE_DEV(gfeatR);
    if (ws == nullptr || ws_bytes < side_voxel_volume_ws_bytes(B, C, H, W) || !is_device_ptr(ws)) {
        set_error("side_voxel_volume_bwd: needs side_voxel_volume_ws_bytes(...) bytes of device workspace");
        return SIDE_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t fsz = (size_t)B * C * H * W;
    float *gl = reinterpret_cast<float *>(ws), *gr = gl + fsz;
    SIDE_CUDA(cudaMemsetAsync(gl, 0, sizeof(float) * 2 * fsz, st));
    const size_t smem = sizeof(float) * 2 * 64 * (kVxTile + 1);
    int rc;
    if ((rc = set_smem_attr((const void *)voxel_volume_bwd_kernel, smem))) return rc;
    voxel_volume_bwd_kernel<<<dim3((kVxN + kVxTile - 1) / kVxTile, N), kVxThreads, smem, st>>>(p, gvoxel, gl, gr);
    SIDE_LAUNCH_CHECK("voxel_volume_bwd_kernel");
    if ((rc = launch_nhwc_to_nchw(gl, gfeatL, B, C, H * W, st))) return rc;
    if ((rc = launch_nhwc_to_nchw(gr, gfeatR, B, C, H * W, st))) return rc;
    return SIDE_OK;
}
