// preprocess.cu -- detector input preparation on the GPU.
//
// Reference: stereoDetector.pre_process (src/lib/modules/stereoDetector.py:45-82): per image, on the host,
//   inp = cv2.warpAffine(image, trans_input, (inp_w, inp_h), flags=cv2.INTER_LINEAR)       uint8 HWC -> uint8 HWC
//   inp = ((inp.astype(float32) / 255.) - mean) / std;  inp.transpose(2, 0, 1)[None]        -> float32 1 x 3 x H x W
// followed by torch.from_numpy and the upload.  Here the raw uint8 image is uploaded once and one kernel does the warp,
// the normalisation and the HWC -> CHW transpose for the left and the right image together.
//
// cv2 is a third-party dependency that is NOT in this image (no opencv wheel): PARITY UNPINNED.  The kernel restates
// OpenCV's published warpAffine algorithm for 8-bit images (imgwarp.cpp, WarpAffineInvoker + remapBilinear):
//   * coordinates in fixed point: AB_BITS = 10, INTER_BITS = 5; for destination (x, y) with the inverse map M,
//       X = saturate(int)(round(M00 x 1024)) + saturate(int)(round((M01 y + M02) 1024)) + 16,  sx = X >> 10, ax = (X >> 5) & 31
//   * bilinear weights from the 32 x 32 table: w = (32 - ay or ay)(32 - ax or ax) * 32  (INTER_REMAP_COEF_SCALE = 2^15; for
//     the linear kernel the products are exact, so the table's sum correction never fires)
//   * value = (sum_taps w * pixel + 2^14) >> 15, taps outside the source are the constant border 0
// oracle/torch_port.py:warp_affine_u8 is the numpy restatement the tests compare with (bit-exact on the uint8 stage).
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace side {

struct PreParams {
    const unsigned char *img[2];      // left, right (right may be NULL)
    float *out[2];
    int sh, sw, dh, dw;
    double m[6];                      // inverse map: dst (x, y) -> src
    float mean[3], sd[3];
};

__device__ __forceinline__ int pre_round_sat(double v)
{
    // cv::saturate_cast<int>(double) = cvRound with saturation: round half to even (lrint)
    v = rint(v);
    return v >= 2147483647.0 ? 2147483647 : (v <= -2147483648.0 ? (int)0x80000000 : (int)v);
}

__global__ void preprocess_kernel(PreParams p)
{
    const int which = blockIdx.z;
    const unsigned char *src = p.img[which];
    float *out = p.out[which];
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= p.dw || src == nullptr) return;
    const int X0 = pre_round_sat((p.m[1] * y + p.m[2]) * 1024.0) + 16, Y0 = pre_round_sat((p.m[4] * y + p.m[5]) * 1024.0) + 16;
    const int X = (X0 + pre_round_sat(p.m[0] * x * 1024.0)) >> 5, Y = (Y0 + pre_round_sat(p.m[3] * x * 1024.0)) >> 5;
    const int sx = X >> 5, sy = Y >> 5, ax = X & 31, ay = Y & 31;
    const int w00 = (32 - ay) * (32 - ax), w01 = (32 - ay) * ax, w10 = ay * (32 - ax), w11 = ay * ax;      // x 32 = table entry
    const bool x0 = sx >= 0 && sx < p.sw, x1 = sx + 1 >= 0 && sx + 1 < p.sw, y0 = sy >= 0 && sy < p.sh, y1 = sy + 1 >= 0 && sy + 1 < p.sh;
    const size_t plane = (size_t)p.dh * p.dw;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int p00 = (x0 && y0) ? src[((size_t)sy * p.sw + sx) * 3 + c] : 0;
        const int p01 = (x1 && y0) ? src[((size_t)sy * p.sw + sx + 1) * 3 + c] : 0;
        const int p10 = (x0 && y1) ? src[((size_t)(sy + 1) * p.sw + sx) * 3 + c] : 0;
        const int p11 = (x1 && y1) ? src[((size_t)(sy + 1) * p.sw + sx + 1) * 3 + c] : 0;
        const int v = ((p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11) * 32 + (1 << 14)) >> 15;      // uint8 result of warpAffine
        const float f = __fdiv_rn((float)min(max(v, 0), 255), 255.f);
        out[(size_t)c * plane + (size_t)y * p.dw + x] = __fdiv_rn(__fsub_rn(f, p.mean[c]), p.sd[c]);
    }
}

}  // namespace side

using namespace side;

extern "C" int side_preprocess_u8(const unsigned char *img_left, const unsigned char *img_right, int src_h, int src_w,
                                  const double *inv_map6, const float *mean3, const float *std3, float *out_left, float *out_right,
                                  int dst_h, int dst_w, void *stream)
{
    SIDE_REQUIRE_DEV(img_left);
    SIDE_REQUIRE_DEV(out_left);
    SIDE_REQUIRE(inv_map6 && mean3 && std3, "side_preprocess_u8: null host argument");
    SIDE_REQUIRE(src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0 && dst_h <= 65535, "side_preprocess_u8: bad shape");
    if (img_right) {
        SIDE_REQUIRE_DEV(img_right);
        SIDE_REQUIRE_DEV(out_right);
    }
    PreParams p;
    p.img[0] = img_left; p.img[1] = img_right;
    p.out[0] = out_left; p.out[1] = out_right;
    p.sh = src_h; p.sw = src_w; p.dh = dst_h; p.dw = dst_w;
    for (int i = 0; i < 6; ++i) p.m[i] = inv_map6[i];
    for (int i = 0; i < 3; ++i) { p.mean[i] = mean3[i]; p.sd[i] = std3[i]; }
    const dim3 grid((unsigned)ceil_div(dst_w, 128), (unsigned)dst_h, img_right ? 2u : 1u);
    preprocess_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    SIDE_LAUNCH_CHECK("preprocess_kernel");
    return SIDE_OK;
}
