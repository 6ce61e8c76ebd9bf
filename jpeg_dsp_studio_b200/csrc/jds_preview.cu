// jds_preview.cu - the GUI's preview downscale (gui/compression_tab.py:532-552):
// cv2.resize(uint8 RGB, (new_w, new_h), interpolation=cv2.INTER_AREA), the step in front of
// the round trip when "preview mode" is on.  Restates OpenCV 4.13 imgproc/resize.cpp:
//   * both shrink factors integers (4K -> 1080p / 720p / 540p): resizeAreaFast_ - integer
//     sum of the fx x fy cell; 2x2 cells -> (sum + 2) >> 2, otherwise
//     saturate_cast<uchar>(sum * (1.f / (fx*fy)))  (float product, round half to even)
//   * otherwise ResizeArea_Invoker<uchar, float>: per source row buf = buf + S*alpha over
//     the x taps, per destination row sum = beta*buf | sum + beta*buf, un-fused fp32, then
//     saturate_cast<uchar>(sum); taps from computeResizeAreaTab (area_span, jds_stages.cuh)
// One thread per destination pixel (3 channels); a warp reads contiguous source bytes.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#include "jds_kernels.cuh"
#include "jds_stages.cuh"

namespace jds {

__device__ __forceinline__ uint8_t sat_round_u8(float v) {
    const int r = __float2int_rn(v);                 // cvRound: round half to even
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

__global__ void __launch_bounds__(256)
k_area_fast_u8(int H, int W, int dh, int dw, int fx, int fy, const uint8_t* __restrict__ src,
               uint8_t* __restrict__ dst) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    int s0 = 0, s1 = 0, s2 = 0;
    for (int j = 0; j < fy; ++j) {
        const uint8_t* p = src + ((size_t)(dy * fy + j) * W + (size_t)dx * fx) * 3;
        for (int i = 0; i < fx; ++i) {
            s0 += p[3 * i];
            s1 += p[3 * i + 1];
            s2 += p[3 * i + 2];
        }
    }
    uint8_t* o = dst + ((size_t)dy * dw + dx) * 3;
    if (fx == 2 && fy == 2) {
        o[0] = (uint8_t)((s0 + 2) >> 2);
        o[1] = (uint8_t)((s1 + 2) >> 2);
        o[2] = (uint8_t)((s2 + 2) >> 2);
    } else {
        const float scale = __fdiv_rn(1.0f, (float)(fx * fy));
        o[0] = sat_round_u8(__fmul_rn((float)s0, scale));
        o[1] = sat_round_u8(__fmul_rn((float)s1, scale));
        o[2] = sat_round_u8(__fmul_rn((float)s2, scale));
    }
}

__global__ void __launch_bounds__(256)
k_area_general_u8(int H, int W, int dh, int dw, const uint8_t* __restrict__ src,
                  uint8_t* __restrict__ dst) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    AreaSpan tx, ty;
    area_span(W, dw, dx, tx);
    area_span(H, dh, dy, ty);
    float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f;
    for (int j = 0; j < ty.n; ++j) {
        const uint8_t* p = src + ((size_t)(ty.s0 + j) * W + tx.s0) * 3;
        float b0 = 0.f, b1 = 0.f, b2 = 0.f;
        for (int i = 0; i < tx.n; ++i) {
            const float a = tx.weight(i);
            b0 = __fadd_rn(b0, __fmul_rn((float)p[3 * i], a));
            b1 = __fadd_rn(b1, __fmul_rn((float)p[3 * i + 1], a));
            b2 = __fadd_rn(b2, __fmul_rn((float)p[3 * i + 2], a));
        }
        const float beta = ty.weight(j);
        if (j == 0) {
            sum0 = __fmul_rn(beta, b0);
            sum1 = __fmul_rn(beta, b1);
            sum2 = __fmul_rn(beta, b2);
        } else {
            sum0 = __fadd_rn(sum0, __fmul_rn(beta, b0));
            sum1 = __fadd_rn(sum1, __fmul_rn(beta, b1));
            sum2 = __fadd_rn(sum2, __fmul_rn(beta, b2));
        }
    }
    uint8_t* o = dst + ((size_t)dy * dw + dx) * 3;
    o[0] = sat_round_u8(sum0);
    o[1] = sat_round_u8(sum1);
    o[2] = sat_round_u8(sum2);
}

// 0 ok, 1 not a shrink (INTER_AREA enlargement is a different OpenCV path: not restated)
int launch_resize_area_u8(int H, int W, int dh, int dw, const uint8_t* src, uint8_t* dst,
                          cudaStream_t s) {
    if (dh < 1 || dw < 1 || dh > H || dw > W) return 1;
    const double sx = (double)W / dw, sy = (double)H / dh;
    const int ix = (int)lrint(sx), iy = (int)lrint(sy);       // saturate_cast<int>(double)
    const bool fast = fabs(sx - ix) < DBL_EPSILON && fabs(sy - iy) < DBL_EPSILON;
    dim3 blk(32, 8), grid((dw + 31) / 32, (dh + 7) / 8);
    if (fast)
        k_area_fast_u8<<<grid, blk, 0, s>>>(H, W, dh, dw, ix, iy, src, dst);
    else
        k_area_general_u8<<<grid, blk, 0, s>>>(H, W, dh, dw, src, dst);
    return 0;
}

}  // namespace jds
