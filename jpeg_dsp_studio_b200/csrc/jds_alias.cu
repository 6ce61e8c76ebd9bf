// jds_alias.cu - kernels of the chroma-aliasing demo's front end (SURVEY 8f #4;
// gui/dialogs/aliasing_demo_dialog.py:125-166): float32 YCrCb conversion, optional 5x5
// Gaussian prefilter evaluated only at the samples the [::2, ::2] decimation keeps, bilinear
// re-enlargement + inverse conversion + clip/truncate to the uint8 frame the hot path then
// compresses at 4:4:4, the integer luma planes compute_metrics compares, and the x10
// difference image.  Per-pixel arithmetic: jds_alias.cuh (bit-identical to OpenCV 4.13).
// HBM-bound elementwise work: one thread per pixel (or decimated sample), coalesced rows.
#include <cuda_runtime.h>
#include <stdint.h>
#include "jds_alias.cuh"
#include "jds_kernels.cuh"

namespace jds {

// RGB uint8 -> Y, Cr, Cb float32 planes (H x W each)
__global__ void __launch_bounds__(256)
k_alias_planes(int H, int W, const uint8_t* __restrict__ rgb, float* __restrict__ Y,
               float* __restrict__ Cr, float* __restrict__ Cb) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const size_t i = (size_t)y * W + x;
    const uint8_t* p = rgb + i * 3;
    float yy, cr, cb;
    alias_forward_px((float)p[0], (float)p[1], (float)p[2], x < 8 * (W / 8), yy, cr, cb);
    Y[i] = yy;
    Cr[i] = cr;
    Cb[i] = cb;
}

// blurred chroma at the even (y, x) positions only: out[hs x ws] for both planes (z = 0: Cr, 1: Cb)
__global__ void __launch_bounds__(256)
k_alias_blur_decimate(int H, int W, int hs, int ws, const float* __restrict__ Cr,
                      const float* __restrict__ Cb, float* __restrict__ Cr_s, float* __restrict__ Cb_s) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x;
    const int yo = blockIdx.y;
    if (xo >= ws) return;
    const float* src = blockIdx.z ? Cb : Cr;
    float* dst = blockIdx.z ? Cb_s : Cr_s;
    dst[(size_t)yo * ws + xo] = alias_blur_at(src, H, W, 2 * xo, 2 * yo);
}

// enlarge the decimated chroma, convert back, clip, truncate
__global__ void __launch_bounds__(256)
k_alias_compose(int H, int W, int hs, int ws, const float* __restrict__ Y,
                const float* __restrict__ Cr_s, const float* __restrict__ Cb_s, size_t row_stride,
                int col_stride, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const float cr = alias_upsample_at(Cr_s, row_stride, col_stride, hs, ws, H, W, x, y);
    const float cb = alias_upsample_at(Cb_s, row_stride, col_stride, hs, ws, H, W, x, y);
    const size_t i = (size_t)y * W + x;
    alias_inverse_px(Y[i], cr, cb, out + i * 3);
}

// integer luma of a uint8 RGB frame, replicated to three channels so that the RGB comparison
// kernels (squared error, SSIM) can be reused: every channel then carries the Y statistics
__global__ void __launch_bounds__(256)
k_alias_luma3(size_t n_px, const uint8_t* __restrict__ rgb, uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const uint8_t* p = rgb + i * 3;
    const uint8_t v = alias_luma_u8(p[0], p[1], p[2]);
    out[i * 3] = out[i * 3 + 1] = out[i * 3 + 2] = v;
}

__global__ void __launch_bounds__(256)
k_alias_diff(size_t n, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
             uint8_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = alias_diff_u8(a[i], b[i]);
}

// scratch: Y, Cr, Cb (H*W floats each) + two decimated planes (hs*ws each)
size_t alias_scratch_floats(int H, int W) {
    const size_t hs = (size_t)(H + 1) / 2, ws = (size_t)(W + 1) / 2;
    return 3 * (size_t)H * W + 2 * hs * ws;
}

int launch_alias_subsample(int H, int W, int prefilter, const uint8_t* rgb, float* scratch,
                           uint8_t* out, cudaStream_t s) {
    const int hs = (H + 1) / 2, ws = (W + 1) / 2;
    const size_t n = (size_t)H * W;
    float *Y = scratch, *Cr = Y + n, *Cb = Cr + n, *Cr_s = Cb + n, *Cb_s = Cr_s + (size_t)hs * ws;
    const dim3 blk(256);
    k_alias_planes<<<dim3((W + 255) / 256, H), blk, 0, s>>>(H, W, rgb, Y, Cr, Cb);
    int launches = 1;
    if (prefilter) {
        k_alias_blur_decimate<<<dim3((ws + 255) / 256, hs, 2), blk, 0, s>>>(H, W, hs, ws, Cr, Cb, Cr_s, Cb_s);
        k_alias_compose<<<dim3((W + 255) / 256, H), blk, 0, s>>>(H, W, hs, ws, Y, Cr_s, Cb_s, (size_t)ws, 1, out);
        launches += 2;
    } else {
        // [::2, ::2] of the full-resolution planes: read them in place with strides
        k_alias_compose<<<dim3((W + 255) / 256, H), blk, 0, s>>>(H, W, hs, ws, Y, Cr, Cb, 2 * (size_t)W, 2, out);
        launches += 1;
    }
    return launches;
}

void launch_alias_luma3(size_t n_px, const uint8_t* rgb, uint8_t* out, cudaStream_t s) {
    k_alias_luma3<<<(unsigned)((n_px + 255) / 256), 256, 0, s>>>(n_px, rgb, out);
}

void launch_alias_diff(size_t n, const uint8_t* a, const uint8_t* b, uint8_t* out, cudaStream_t s) {
    k_alias_diff<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, a, b, out);
}

}  // namespace jds
