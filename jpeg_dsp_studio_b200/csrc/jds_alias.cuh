// jds_alias.cuh - per-work-item arithmetic of the reference's chroma-aliasing demo
// (gui/dialogs/aliasing_demo_dialog.py:125-150): the OpenCV FLOAT32 kernels
// the demo runs in front of the hot path, in OpenCV 4.13's operation order (probed against
// the library and pinned by oracle/aliasing_port.py rules F1-F5):
//   F1 cvtColor(RGB2YCrCb) on a float image   F2 GaussianBlur(5x5, 0.8)
//   F3 resize(INTER_LINEAR)                   F4 cvtColor(YCrCb2RGB)
//   F5 cvtColor(RGB2YCrCb) on uint8 (the Y channel compute_metrics reads, :69-83)
// Every float operation is individually rounded; an FMA appears only where OpenCV's
// vectorised code has one.  __host__ __device__ so tests/emul runs the same source on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include "jds_math.cuh"

namespace jds {

struct F32 {
#if defined(__CUDA_ARCH__)
    static JDS_HD float add(float a, float b) { return __fadd_rn(a, b); }
    static JDS_HD float sub(float a, float b) { return __fsub_rn(a, b); }
    static JDS_HD float mul(float a, float b) { return __fmul_rn(a, b); }
    static JDS_HD float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static JDS_HD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static JDS_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
#else
    static JDS_HD float add(float a, float b) { volatile float r = a + b; return r; }
    static JDS_HD float sub(float a, float b) { volatile float r = a - b; return r; }
    static JDS_HD float mul(float a, float b) { volatile float r = a * b; return r; }
    static JDS_HD float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
    static JDS_HD double dfma(double a, double b, double c) { return ::fma(a, b, c); }
    static JDS_HD double ddiv(double a, double b) { volatile double r = a / b; return r; }
#endif
};

// cv2.getGaussianKernel(5, 0.8, CV_32F)
#define JDS_AK0 0x1.674b98p-6f
#define JDS_AK1 0x1.d3fe2ep-3f
#define JDS_AK2 0x1.ff1860p-2f

JDS_HD int alias_reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// F1: one pixel at column x of a W-wide row.  Columns below 8*floor(W/8) go through the
// vector body, the rest through the scalar tail (different association of the Y sum).
JDS_HD void alias_forward_px(float R, float G, float B, bool body, float& Y, float& Cr, float& Cb) {
    const float c0 = 0.299f, c1 = 0.587f, c2 = 0.114f;
    Y = body ? F32::fma(R, c0, F32::fma(G, c1, F32::mul(B, c2)))
             : F32::fma(B, c2, F32::fma(R, c0, F32::mul(G, c1)));
    Cr = F32::fma(F32::sub(R, Y), 0.713f, 0.5f);
    Cb = F32::fma(F32::sub(B, Y), 0.564f, 0.5f);
}

// F2 row pass at (row pointer, x); `plain` = last column of an odd width
JDS_HD float alias_blur_row(const float* row, int x, int W, bool plain) {
    const float m2 = row[alias_reflect101(x - 2, W)], m1 = row[alias_reflect101(x - 1, W)];
    const float c = row[x];
    const float p1 = row[alias_reflect101(x + 1, W)], p2 = row[alias_reflect101(x + 2, W)];
    const float s2 = F32::add(m2, p2), s1 = F32::add(m1, p1);
    if (plain)
        return F32::add(F32::add(F32::mul(c, JDS_AK2), F32::mul(s1, JDS_AK1)), F32::mul(s2, JDS_AK0));
    return F32::fma(s2, JDS_AK0, F32::fma(c, JDS_AK2, F32::mul(s1, JDS_AK1)));
}

// F2 at one sample (x, y) of an H x W plane: five row-filtered values, then the column pass
JDS_HD float alias_blur_at(const float* plane, int H, int W, int x, int y) {
    const bool row_plain = (W & 1) && x == W - 1;
    const bool col_fma = x < 8 * (W / 8);
    float r[5];
    for (int d = -2; d <= 2; ++d)
        r[d + 2] = alias_blur_row(plane + (size_t)alias_reflect101(y + d, H) * W, x, W, row_plain);
    const float s2 = F32::add(r[0], r[4]), s1 = F32::add(r[1], r[3]);
    if (col_fma) return F32::fma(s2, JDS_AK0, F32::fma(s1, JDS_AK1, F32::mul(r[2], JDS_AK2)));
    return F32::add(F32::add(F32::mul(r[2], JDS_AK2), F32::mul(s1, JDS_AK1)), F32::mul(s2, JDS_AK0));
}

// F3 taps of destination index d (n_dst samples from n_src): source coordinate in fp64 with
// one FMA, weight cast to float32
JDS_HD void alias_linear_tap(int d, int n_dst, int n_src, int& i0, int& i1, float& w) {
    const double f = F32::dfma((double)d + 0.5, F32::ddiv((double)n_src, (double)n_dst), -0.5);
    const double s = floor(f);
    w = (float)(f - s);
    const int si = (int)s;
    i0 = si < 0 ? 0 : (si > n_src - 1 ? n_src - 1 : si);
    i1 = si + 1 < 0 ? 0 : (si + 1 > n_src - 1 ? n_src - 1 : si + 1);
}

// F3 at destination (x, y): `src` is the decimated plane, sample (i, j) at src[j*row_stride + i*col_stride]
JDS_HD float alias_upsample_at(const float* src, size_t row_stride, int col_stride, int hs, int ws,
                               int H, int W, int x, int y) {
    int x0, x1, y0, y1;
    float wx, wy;
    alias_linear_tap(x, W, ws, x0, x1, wx);
    alias_linear_tap(y, H, hs, y0, y1, wy);
    const float* r0 = src + (size_t)y0 * row_stride;
    const float* r1 = src + (size_t)y1 * row_stride;
    const float a0 = r0[(size_t)x0 * col_stride], b0 = r0[(size_t)x1 * col_stride];
    const float a1 = r1[(size_t)x0 * col_stride], b1 = r1[(size_t)x1 * col_stride];
    const float t0 = F32::fma(F32::sub(b0, a0), wx, a0);
    const float t1 = F32::fma(F32::sub(b1, a1), wx, a1);
    return F32::fma(F32::sub(t1, t0), wy, t0);
}

// np.clip(v, 0, 255).astype(np.uint8): clamp, then truncate toward zero
JDS_HD uint8_t alias_clip_u8(float v) {
    v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
    return (uint8_t)(int)v;
}

// F4 + clip + truncate
JDS_HD void alias_inverse_px(float Y, float Cr, float Cb, uint8_t* rgb) {
    const float cr = F32::sub(Cr, 0.5f), cb = F32::sub(Cb, 0.5f);
    rgb[0] = alias_clip_u8(F32::fma(cr, 1.403f, Y));
    rgb[1] = alias_clip_u8(F32::fma(cr, -0.714f, F32::fma(cb, -0.344f, Y)));
    rgb[2] = alias_clip_u8(F32::fma(cb, 1.773f, Y));
}

// F5: Y of cvtColor(uint8 RGB, COLOR_RGB2YCrCb)
JDS_HD uint8_t alias_luma_u8(int r, int g, int b) {
    return (uint8_t)((r * 4899 + g * 9617 + b * 1868 + (1 << 13)) >> 14);
}

// _compute_difference (:162-166): clip(|a - b| * 10, 0, 255) truncated; exact in integers
JDS_HD uint8_t alias_diff_u8(int a, int b) {
    const int d = (a > b ? a - b : b - a) * 10;
    return (uint8_t)(d > 255 ? 255 : d);
}

}  // namespace jds
