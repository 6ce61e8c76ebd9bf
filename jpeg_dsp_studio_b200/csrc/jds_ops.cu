// jds_ops.cu - the reference's stage functions as stand-alone device operators, exact
// (fp64, reference op order) arithmetic:
//   engines/color_space.py:8-24   rgb_to_ycbcr / ycbcr_to_rgb          k_color_f64
//   engines/color_space.py:27-53  subsample_chroma (one plane)         k_subsample_plane
//   engines/color_space.py:56-66  upsample_chroma (one plane)          k_upsample_plane
//   utils/metrics.py:9-28         compute_psnr_ssim partial sums       k_sse_u8 (+ SSIM kernels)
//   utils/metrics.py:63-83        estimate_bitrate_no_entropy counts   k_bitcount
// The fused / staged round-trip kernels contain the same arithmetic; these exist so that each
// row of the path can be called and checked on its own (tests/test_stage_ops_gpu.py).
#include <cuda_runtime.h>
#include <stdint.h>
#include "jds_kernels.cuh"
#include "jds_stages.cuh"

namespace jds {

__global__ void __launch_bounds__(256)
k_color_f64(int direction, long long n, const double* __restrict__ in, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = in[3 * i], b = in[3 * i + 1], c = in[3 * i + 2];
    double x, y, z;
    if (direction == 0) rgb_to_ycbcr<Exact>(a, b, c, x, y, z);
    else ycbcr_to_rgb<Exact>(a, b, c, x, y, z);          // includes np.clip(0, 255)
    out[3 * i] = x;
    out[3 * i + 1] = y;
    out[3 * i + 2] = z;
}

// plane sample after the optional cv2.GaussianBlur(3x3, 0.75) (A2)
template <bool PF>
__device__ __forceinline__ double plane_at(const double* __restrict__ p, int H, int W, int W4,
                                           int y, int x) {
    if (!PF) return p[(size_t)y * W + x];
    double r[3];
    const bool tail = x >= W4;
    const int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double* row = p + (size_t)reflect101(y - 1 + i, H) * W;
        r[i] = blur_row<Exact>(row[xm], row[x], row[xp], tail);
    }
    return blur_col<Exact>(r[0], r[1], r[2]);
}

template <bool PF>
__global__ void __launch_bounds__(256)
k_subsample_plane(const double* __restrict__ p, int H, int W, int hc, int wc, int sub,
                  int general, double* __restrict__ out) {
    typedef Exact E;
    const int cx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cy = blockIdx.y * blockDim.y + threadIdx.y;
    if (cx >= wc || cy >= hc) return;
    const int W4 = 4 * (W / 4);
    double v;
    if (general) {
        AreaSpan tx, ty;
        area_span(W, wc, cx, tx);
        area_span(H, hc, cy, ty);
        v = 0.0;
        for (int j = 0; j < ty.n; ++j) {
            double buf = 0.0;
            for (int i = 0; i < tx.n; ++i)
                buf = E::add(buf, E::mul(plane_at<PF>(p, H, W, W4, ty.s0 + j, tx.s0 + i),
                                         (double)tx.weight(i)));
            const double beta = (double)ty.weight(j);
            v = j == 0 ? E::mul(beta, buf) : E::add(v, E::mul(beta, buf));
        }
    } else if (sub == 1) {                                   // (a + b) * 0.5
        v = E::mul(E::add(plane_at<PF>(p, H, W, W4, cy, 2 * cx),
                          plane_at<PF>(p, H, W, W4, cy, 2 * cx + 1)), 0.5);
    } else {                                                 // (((a + b) + c) + d) * 0.25
        const double a = plane_at<PF>(p, H, W, W4, 2 * cy, 2 * cx);
        const double b = plane_at<PF>(p, H, W, W4, 2 * cy, 2 * cx + 1);
        const double c = plane_at<PF>(p, H, W, W4, 2 * cy + 1, 2 * cx);
        const double d = plane_at<PF>(p, H, W, W4, 2 * cy + 1, 2 * cx + 1);
        v = E::mul(E::add(E::add(E::add(a, b), c), d), 0.25);
    }
    out[(size_t)cy * wc + cx] = v;
}

__global__ void __launch_bounds__(256)
k_upsample_plane(const double* __restrict__ p, int h, int w, int H, int W,
                 double* __restrict__ out) {
    typedef Exact E;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    int x0 = x, x1 = x, y0 = y, y1 = y;
    double fx = 0.0, fy = 0.0;
    if (w != W) upsample_taps<E>(x, w, (double)w / (double)W, x0, x1, fx);
    if (h != H) upsample_taps<E>(y, h, (double)h / (double)H, y0, y1, fy);
    // horizontal lerp on the two source rows, then vertical (IPP order, A8)
    const double* r0 = p + (size_t)y0 * w;
    double t0 = r0[x0];
    if (w != W) t0 = E::fma(E::sub(r0[x1], r0[x0]), fx, r0[x0]);
    double v = t0;
    if (h != H) {
        const double* r1 = p + (size_t)y1 * w;
        double t1 = r1[x0];
        if (w != W) t1 = E::fma(E::sub(r1[x1], r1[x0]), fx, r1[x0]);
        v = E::fma(E::sub(t1, t0), fy, t0);
    }
    out[(size_t)y * W + x] = v;
}

// squared errors of two uint8 RGB images: integer for R,G,B (utils/metrics.py:11), fp64 for
// BT.601 Y computed from the uint8 values (utils/metrics.py:17-20)
__global__ void __launch_bounds__(256)
k_sse_u8(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, long long n_px,
         DevMetrics* __restrict__ m) {
    unsigned long long s = 0;
    double sy = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_px;
         i += (long long)gridDim.x * blockDim.x) {
        const uint8_t* pa = a + 3 * i;
        const uint8_t* pb = b + 3 * i;
        const int dr = (int)pa[0] - (int)pb[0], dg = (int)pa[1] - (int)pb[1], db = (int)pa[2] - (int)pb[2];
        s += (unsigned)(dr * dr + dg * dg + db * db);
        const double ya = luma601<Exact>((double)pa[0], (double)pa[1], (double)pa[2]);
        const double yb = luma601<Exact>((double)pb[0], (double)pb[1], (double)pb[2]);
        const double d = ya - yb;
        sy += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        sy += __shfl_down_sync(0xffffffffu, sy, o);
    }
    __shared__ unsigned long long sh_s[8];
    __shared__ double sh_y[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh_s[warp] = s; sh_y[warp] = sy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) { s += sh_s[k]; sy += sh_y[k]; }
        atomicAdd(&m->sse_rgb, s);
        atomicAdd(&m->sse_y, sy);
    }
}

// non-zero count and 6 + ceil(log2(|v|+1)) + 1 bits per non-zero coefficient
__global__ void __launch_bounds__(256)
k_bitcount(const int16_t* __restrict__ c, unsigned long long n, DevMetrics* __restrict__ m) {
    unsigned long long bits = 0, nnz = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const int v = c[i];
        bits += (unsigned)coeff_bits(v);
        nnz += v != 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bits += __shfl_down_sync(0xffffffffu, bits, o);
        nnz += __shfl_down_sync(0xffffffffu, nnz, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&m->coeff_bits, bits);
        atomicAdd(&m->nnz, nnz);
    }
}

void launch_color_f64(int direction, long long n, const double* in, double* out, cudaStream_t s) {
    k_color_f64<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(direction, n, in, out);
}

void launch_subsample_plane(const double* p, int H, int W, int hc, int wc, int sub, int prefilter,
                            double* out, cudaStream_t s) {
    const int general = (W % 2) || (sub == 2 && (H % 2));
    dim3 blk(32, 8), grid((wc + 31) / 32, (hc + 7) / 8);
    if (prefilter) k_subsample_plane<true><<<grid, blk, 0, s>>>(p, H, W, hc, wc, sub, general, out);
    else k_subsample_plane<false><<<grid, blk, 0, s>>>(p, H, W, hc, wc, sub, general, out);
}

void launch_upsample_plane(const double* p, int h, int w, int H, int W, double* out, cudaStream_t s) {
    dim3 blk(32, 8), grid((W + 31) / 32, (H + 7) / 8);
    k_upsample_plane<<<grid, blk, 0, s>>>(p, h, w, H, W, out);
}

void launch_sse_u8(const uint8_t* a, const uint8_t* b, long long n_px, DevMetrics* m, int sm_count,
                   cudaStream_t s) {
    long long want = (n_px + 255) / 256;
    unsigned grid = (unsigned)(want < (long long)sm_count * 8 ? (want ? want : 1) : (long long)sm_count * 8);
    k_sse_u8<<<grid, 256, 0, s>>>(a, b, n_px, m);
}

void launch_bitcount(const int16_t* c, unsigned long long n, DevMetrics* m, int sm_count, cudaStream_t s) {
    unsigned long long want = (n + 255) / 256;
    unsigned grid = (unsigned)(want < (unsigned long long)sm_count * 8 ? (want ? want : 1)
                                                                      : (unsigned long long)sm_count * 8);
    k_bitcount<<<grid, 256, 0, s>>>(c, n, m);
}

}  // namespace jds
