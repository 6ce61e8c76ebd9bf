// jds_kernels.cu - staged kernels of the round trip (sm_100a).
//
// Stage 1  k_forward   RGB uint8 -> Y / decimated Cb / Cr planes          (per chroma cell)
// Stage 2  k_codec     8x8 block DCT, quantise, bit model, IDCT           (per block)
// Stage 3  k_inverse   upsample, YCbCr->RGB, truncate, SSE, error maps    (per pixel)
// Stage 4  k_ssim      7x7-window SSIM of R, G, B and Y                   (per 32x16 tile)
//
// blockIdx.z is the unit (frame of a batch, or quality point of a sweep).
#include <cuda_runtime.h>
#include "jds_kernels.cuh"
#include "jds_stages.cuh"

namespace jds {

// ------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------
// Stage 1
// ------------------------------------------------------------------------------
template <class P, int SUB, bool PF>
__global__ void __launch_bounds__(256)
k_forward(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
          typename P::T* __restrict__ fwd, size_t fwd_stride, int chroma_only) {
    const int cx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cy = blockIdx.y * blockDim.y + threadIdx.y;
    if (cx >= g.wc || cy >= g.hc) return;
    const uint8_t* in = rgb + (size_t)blockIdx.z * rgb_stride;
    typename P::T* Y = fwd + (size_t)blockIdx.z * fwd_stride;
    typename P::T* Cb = Y + g.plane_y;
    typename P::T* Cr = Cb + g.plane_c;
    forward_cell<P, SUB, PF>(g, in, cx, cy, chroma_only ? nullptr : Y, Cb, Cr);
}

// general geometry (odd width, or odd height under 4:2:0): one thread per luma pixel; the
// threads inside the chroma plane also produce one area-weighted chroma sample each
template <class P, bool PF>
__global__ void __launch_bounds__(256)
k_forward_general(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
                  typename P::T* __restrict__ fwd, size_t fwd_stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const uint8_t* in = rgb + (size_t)blockIdx.z * rgb_stride;
    typename P::T* Y = fwd + (size_t)blockIdx.z * fwd_stride;
    typename P::T* Cb = Y + g.plane_y;
    typename P::T* Cr = Cb + g.plane_c;
    forward_luma<P>(g, in, x, y, Y);
    if (x < g.wc && y < g.hc) forward_chroma_area<P, PF>(g, in, x, y, Cb, Cr);
}

// ------------------------------------------------------------------------------
// Stage 2
// ------------------------------------------------------------------------------
template <class P, bool HIST>
__global__ void __launch_bounds__(128)
k_codec(Geom g, const typename P::T* __restrict__ fwd, size_t fwd_stride,
        typename P::T* __restrict__ rec, size_t rec_stride,
        const QTables* __restrict__ tables, int table_stride,
        int16_t* __restrict__ coeffs, size_t coeff_stride,
        DevMetrics* __restrict__ metrics, long long first_block) {
    typedef typename P::T T;
    __shared__ QTables tb;
    __shared__ unsigned int s_hist[50];
    const int unit = blockIdx.z;
    {
        const QTables* src = tables + (size_t)unit * table_stride;
        for (int i = threadIdx.x; i < 64; i += blockDim.x) {
            tb.q[i] = src->q[i];
            tb.rq[i] = src->rq[i];
            tb.dqx[i] = src->dqx[i];
            tb.fq[i] = src->fq[i];
            tb.dq[i] = src->dq[i];
        }
        if (HIST)
            for (int i = threadIdx.x; i < 50; i += blockDim.x) s_hist[i] = 0;
    }
    __syncthreads();

    const long long total = g.nblk_y + 2 * g.nblk_c;
    const long long b = first_block + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bits = 0, nnz = 0;
    if (b < total) {
        // block -> (plane, bx, by); coefficient order is Y | Cb | Cr, block raster
        int plane = 0;
        long long lb = b;
        if (lb >= g.nblk_y) {
            lb -= g.nblk_y;
            plane = 1;
            if (lb >= g.nblk_c) {
                lb -= g.nblk_c;
                plane = 2;
            }
        }
        const int nbx = plane ? g.nbx_c : g.nbx_y;
        const int h = plane ? g.hc : g.H, w = plane ? g.wc : g.W;
        const int stride = plane ? g.wcp : g.Wp;
        const size_t poff = plane == 0 ? 0 : (plane == 1 ? g.plane_y : g.plane_y + g.plane_c);
        const int by = (int)(lb / nbx), bx = (int)(lb % nbx);
        const T* src = fwd + (size_t)unit * fwd_stride + poff;
        T* dst = rec + (size_t)unit * rec_stride + poff;

        T v[64];
        load_block<T>(src, stride, h, w, bx, by, v);
        int16_t q[64];
        BlockStats st;
        BlockCodec<P>::run(v, q, tb, st, nullptr, nullptr);
        store_block<T>(dst, stride, h, w, bx, by, v);
        bits = st.bits;
        nnz = st.nnz;
        if (coeffs) {
            uint4* out = reinterpret_cast<uint4*>(coeffs + (size_t)unit * coeff_stride + (size_t)b * 64);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint4 pk;
                pk.x = (uint16_t)q[i * 8 + 0] | ((uint32_t)(uint16_t)q[i * 8 + 1] << 16);
                pk.y = (uint16_t)q[i * 8 + 2] | ((uint32_t)(uint16_t)q[i * 8 + 3] << 16);
                pk.z = (uint16_t)q[i * 8 + 4] | ((uint32_t)(uint16_t)q[i * 8 + 5] << 16);
                pk.w = (uint16_t)q[i * 8 + 6] | ((uint32_t)(uint16_t)q[i * 8 + 7] << 16);
                out[i] = pk;
            }
        }
        if (HIST) {
#pragma unroll 8
            for (int i = 0; i < 64; ++i) {
                const int hb = hist_bin((int)q[i]);
                if (hb >= 0) atomicAdd(&s_hist[hb], 1u);
            }
        }
    }
    bits = warp_sum_u64(bits);
    nnz = warp_sum_u64(nnz);
    DevMetrics* m = metrics + unit;
    if ((threadIdx.x & 31) == 0) {
        if (bits) atomicAdd(&m->coeff_bits, bits);
        if (nnz) atomicAdd(&m->nnz, nnz);
    }
    if (HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < 50; i += blockDim.x)
            if (s_hist[i]) atomicAdd(&m->hist[i], (unsigned long long)s_hist[i]);
    }
}

// ------------------------------------------------------------------------------
// Stage 3
// ------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(256)
k_inverse(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
          const typename P::T* __restrict__ fwd, size_t fwd_stride,
          const typename P::T* __restrict__ rec, size_t rec_stride,
          uint8_t* __restrict__ recon, size_t recon_stride,
          double* __restrict__ err_y, double* __restrict__ err_rgb,
          DevMetrics* __restrict__ metrics) {
    typedef typename P::T T;
    const int unit = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    unsigned long long sse = 0;
    double ssey = 0.0;
    if (x < g.W && y < g.H) {
        const uint8_t* in = rgb + (size_t)unit * rgb_stride;
        const T* Yf = fwd + (size_t)unit * fwd_stride;
        const T* Yr = rec + (size_t)unit * rec_stride;
        const T* Cbr = Yr + g.plane_y;
        const T* Crr = Cbr + g.plane_c;
        PixelOut o = inverse_pixel<P>(g, in, x, y, Yf, Yr, Cbr, Crr);
        uint8_t* out = recon + (size_t)unit * recon_stride + ((size_t)y * g.W + x) * 3;
        out[0] = o.r;
        out[1] = o.g;
        out[2] = o.b;
        if (err_y) err_y[(size_t)y * g.W + x] = o.err_y;
        if (err_rgb) err_rgb[(size_t)y * g.W + x] = o.err_rgb;
        sse = o.sse_rgb;
        ssey = o.sse_y;
    }
    __shared__ unsigned long long s_sse[8];
    __shared__ double s_ssey[8];
    sse = warp_sum_u64(sse);
    ssey = warp_sum_f64(ssey);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if ((tid & 31) == 0) {
        s_sse[tid >> 5] = sse;
        s_ssey[tid >> 5] = ssey;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long a = 0;
        double bsum = 0.0;
        const int nw = (blockDim.x * blockDim.y) >> 5;
        for (int i = 0; i < nw; ++i) {
            a += s_sse[i];
            bsum += s_ssey[i];
        }
        DevMetrics* m = metrics + unit;
        if (a) atomicAdd(&m->sse_rgb, a);
        if (bsum != 0.0) atomicAdd(&m->sse_y, bsum);
    }
}

// Stage 3, four pixels per thread (W % 4 == 0): same per-pixel arithmetic, but the 12 output
// bytes leave as three 32-bit stores and the error maps as 16-byte stores.
template <class P>
__global__ void __launch_bounds__(256)
k_inverse4(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
           const typename P::T* __restrict__ fwd, size_t fwd_stride,
           const typename P::T* __restrict__ rec, size_t rec_stride,
           uint8_t* __restrict__ recon, size_t recon_stride,
           double* __restrict__ err_y, double* __restrict__ err_rgb,
           DevMetrics* __restrict__ metrics) {
    typedef typename P::T T;
    const int unit = blockIdx.z;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    unsigned long long sse = 0;
    double ssey = 0.0;
    if (x < g.W && y < g.H) {
        const uint8_t* in = rgb + (size_t)unit * rgb_stride;
        const T* Yf = fwd + (size_t)unit * fwd_stride;
        const T* Yr = rec + (size_t)unit * rec_stride;
        const T* Cbr = Yr + g.plane_y;
        const T* Crr = Cbr + g.plane_c;
        uint32_t by[12];
        double ey[4], ergb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const PixelOut o = inverse_pixel<P>(g, in, x + i, y, Yf, Yr, Cbr, Crr);
            by[3 * i] = o.r; by[3 * i + 1] = o.g; by[3 * i + 2] = o.b;
            ey[i] = o.err_y;
            ergb[i] = o.err_rgb;
            sse += o.sse_rgb;
            ssey += o.sse_y;
        }
        uint32_t* out = reinterpret_cast<uint32_t*>(recon + (size_t)unit * recon_stride +
                                                    ((size_t)y * g.W + x) * 3);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            out[i] = by[4 * i] | (by[4 * i + 1] << 8) | (by[4 * i + 2] << 16) | (by[4 * i + 3] << 24);
        if (err_y) {
            double2* d = reinterpret_cast<double2*>(err_y + (size_t)y * g.W + x);
            d[0] = make_double2(ey[0], ey[1]);
            d[1] = make_double2(ey[2], ey[3]);
        }
        if (err_rgb) {
            double2* d = reinterpret_cast<double2*>(err_rgb + (size_t)y * g.W + x);
            d[0] = make_double2(ergb[0], ergb[1]);
            d[1] = make_double2(ergb[2], ergb[3]);
        }
    }
    __shared__ unsigned long long s_sse[8];
    __shared__ double s_ssey[8];
    sse = warp_sum_u64(sse);
    ssey = warp_sum_f64(ssey);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if ((tid & 31) == 0) {
        s_sse[tid >> 5] = sse;
        s_ssey[tid >> 5] = ssey;
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long a = 0;
        double bsum = 0.0;
        const int nw = (blockDim.x * blockDim.y) >> 5;
        for (int i = 0; i < nw; ++i) {
            a += s_sse[i];
            bsum += s_ssey[i];
        }
        DevMetrics* m = metrics + unit;
        if (a) atomicAdd(&m->sse_rgb, a);
        if (bsum != 0.0) atomicAdd(&m->sse_y, bsum);
    }
}

// ------------------------------------------------------------------------------
// Stage 4: SSIM.  A CTA owns SS_TW x SS_TH window positions; window (i, j) covers
// pixels [i, i+7) x [j, j+7) (skimage crops the 3-pixel border of the filtered map,
// which leaves exactly the windows that lie inside the image).
// ------------------------------------------------------------------------------
constexpr int SS_TW = 32, SS_TH = 16, SS_IW = SS_TW + 6, SS_IH = SS_TH + 6;

template <class T>
__global__ void __launch_bounds__(256)
k_ssim(int H, int W, const uint8_t* __restrict__ a_img, size_t a_stride,
       const uint8_t* __restrict__ b_img, size_t b_stride, DevMetrics* __restrict__ metrics) {
    __shared__ T s_x[SS_IH][SS_IW];
    __shared__ T s_y[SS_IH][SS_IW];
    __shared__ T s_h[5][SS_IH][SS_TW];
    __shared__ double s_red[8];
    const int unit = blockIdx.z;
    const uint8_t* A = a_img + (size_t)unit * a_stride;
    const uint8_t* B = b_img + (size_t)unit * b_stride;
    const int ox = blockIdx.x * SS_TW, oy = blockIdx.y * SS_TH;
    const int nwx = W - 6, nwy = H - 6;
    const int tid = threadIdx.x;
    for (int ch = 0; ch < 4; ++ch) {
        const T shift = ch < 3 ? T(0) : T(128);
        for (int i = tid; i < SS_IH * SS_IW; i += 256) {
            const int r = i / SS_IW, c = i % SS_IW;
            const int yy = oy + r, xx = ox + c;
            T xv = 0, yv = 0;
            if (yy < H && xx < W) {
                const uint8_t* pa = A + ((size_t)yy * W + xx) * 3;
                const uint8_t* pb = B + ((size_t)yy * W + xx) * 3;
                if (ch < 3) {
                    xv = T(pa[ch]);
                    yv = T(pb[ch]);
                } else {
                    // Y is not integer valued: centre it so the fp32 squares keep their
                    // low bits (variances are shift invariant, the means are un-shifted
                    // inside ssim_from_sums)
                    xv = T(luma601<Exact>((double)pa[0], (double)pa[1], (double)pa[2]) - 128.0);
                    yv = T(luma601<Exact>((double)pb[0], (double)pb[1], (double)pb[2]) - 128.0);
                }
            }
            s_x[r][c] = xv;
            s_y[r][c] = yv;
        }
        __syncthreads();
        for (int i = tid; i < SS_IH * SS_TW; i += 256) {
            const int r = i / SS_TW, c = i % SS_TW;
            T sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const T xv = s_x[r][c + k], yv = s_y[r][c + k];
                sx += xv;
                sy += yv;
                sxx += xv * xv;
                syy += yv * yv;
                sxy += xv * yv;
            }
            s_h[0][r][c] = sx;
            s_h[1][r][c] = sy;
            s_h[2][r][c] = sxx;
            s_h[3][r][c] = syy;
            s_h[4][r][c] = sxy;
        }
        __syncthreads();
        double acc = 0.0;
        for (int i = tid; i < SS_TH * SS_TW; i += 256) {
            const int r = i / SS_TW, c = i % SS_TW;
            if (oy + r < nwy && ox + c < nwx) {
                T m[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    T s = 0;
#pragma unroll
                    for (int k = 0; k < 7; ++k) s += s_h[q][r + k][c];
                    m[q] = s;
                }
                acc += (double)ssim_from_sums<T>(m[0], m[1], m[2], m[3], m[4], shift);
            }
        }
        acc = warp_sum_f64(acc);
        if ((tid & 31) == 0) s_red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int i = 0; i < 8; ++i) t += s_red[i];
            atomicAdd(&metrics[unit].ssim_sum[ch], t);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------
// one selected block (IntermediateData.selected_block_*, pipeline.py:126-151)
// ------------------------------------------------------------------------------
struct SelectedOut {
    double original[64], shifted[64], dct[64], dequantized[64], reconstructed[64];
    int16_t quantized[64];
};

__global__ void k_selected_block(Geom g, const uint8_t* __restrict__ rgb, int bx, int by,
                                 const QTables* __restrict__ tables, SelectedOut* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double v[64];
    for (int r = 0; r < 8; ++r) {
        const int yy = reflect_index(by * 8 + r, g.H);
        for (int c = 0; c < 8; ++c) {
            const int xx = reflect_index(bx * 8 + c, g.W);
            double rr, gg, bb;
            load_rgb<Exact>(rgb, g.W, yy, xx, rr, gg, bb);
            v[r * 8 + c] = luma601<Exact>(rr, gg, bb);
        }
    }
    for (int i = 0; i < 64; ++i) {
        out->original[i] = v[i];
        out->shifted[i] = Exact::sub(v[i], 128.0);
    }
    BlockStats st;
    BlockCodec<Exact>::run(v, out->quantized, *tables, st, out->dct, out->dequantized);
    for (int i = 0; i < 64; ++i) out->reconstructed[i] = v[i];
}

// ------------------------------------------------------------------------------
// 50-bin histogram of the int16 coefficients (np.histogram(all, 50, (-100, 100)),
// engines/pipeline.py:124): one pass over the coefficient buffer, 8 values per 16-byte
// load, per-WARP shared-memory histograms (no cross-warp contention) with equal bins of a
// warp merged by __match_any_sync before the atomic.
// ------------------------------------------------------------------------------
constexpr int HIST_WARPS = 8;

__global__ void __launch_bounds__(HIST_WARPS * 32)
k_hist50(const int16_t* __restrict__ coeffs, size_t coeff_stride, size_t n_coeffs,
         DevMetrics* __restrict__ metrics) {
    __shared__ unsigned int s_h[HIST_WARPS][52];
    const int unit = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < HIST_WARPS * 52; i += blockDim.x) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(coeffs + (size_t)unit * coeff_stride);
    const size_t n_vec = n_coeffs / 8;                       // blocks of 64: always a multiple of 8
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec;
         v += (size_t)gridDim.x * blockDim.x) {
        const uint4 q = __ldg(src + v);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int val = (int)(int16_t)(w[k >> 1] >> ((k & 1) * 16));
            const int b = hist_bin(val);                     // -1: outside [-100, 100]
            const unsigned peers = __match_any_sync(__activemask(), b);
            if (b >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(&s_h[warp][b], (unsigned)__popc(peers));
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 50; b += blockDim.x) {
        unsigned t = 0;
#pragma unroll
        for (int wv = 0; wv < HIST_WARPS; ++wv) t += s_h[wv][b];
        if (t) atomicAdd(&metrics[unit].hist[b], (unsigned long long)t);
    }
}

void launch_hist50(const int16_t* coeffs, size_t coeff_stride, size_t n_coeffs, DevMetrics* metrics,
                   int units, int sm_count, cudaStream_t s) {
    size_t want = (n_coeffs / 8 + HIST_WARPS * 32 - 1) / (HIST_WARPS * 32);
    unsigned gx = (unsigned)(want < (size_t)sm_count * 8 ? (want ? want : 1) : (size_t)sm_count * 8);
    dim3 grid(gx, units);
    k_hist50<<<grid, HIST_WARPS * 32, 0, s>>>(coeffs, coeff_stride, n_coeffs, metrics);
}

// ------------------------------------------------------------------------------
// GUI plot payload: what the reference's plots draw, reduced on the device
//   * value histogram: count of every int16 coefficient value (index v + 1024); the host
//     bins it the way matplotlib's ax.hist / np.histogram(bins=50) does
//     (gui/widgets/mpl_canvas.py:81-100, gui/compression_tab.py:662-667)
//   * heat map: clip(err * 10, 0, 255) (gui/widgets/mpl_canvas.py:116-118) truncated to uint8
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(HIST_WARPS * 32)
k_value_hist(const int16_t* __restrict__ coeffs, size_t n_coeffs,
             unsigned long long* __restrict__ hist) {
    __shared__ unsigned int s_h[VALUE_HIST_BINS];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < VALUE_HIST_BINS; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(coeffs);
    const size_t n_vec = n_coeffs / 8;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec;
         v += (size_t)gridDim.x * blockDim.x) {
        const uint4 q = __ldg(src + v);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int val = (int)(int16_t)(w[k >> 1] >> ((k & 1) * 16));
            const int b = val + VALUE_HIST_BINS / 2;          // |val| <= 1016 by construction
            const unsigned peers = __match_any_sync(__activemask(), b);
            if ((unsigned)b < (unsigned)VALUE_HIST_BINS && lane == (__ffs(peers) - 1))
                atomicAdd(&s_h[b], (unsigned)__popc(peers));
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < VALUE_HIST_BINS; b += blockDim.x)
        if (s_h[b]) atomicAdd(&hist[b], (unsigned long long)s_h[b]);
}

void launch_value_hist(const int16_t* coeffs, size_t n_coeffs, unsigned long long* hist,
                       int sm_count, cudaStream_t s) {
    size_t want = (n_coeffs / 8 + HIST_WARPS * 32 - 1) / (HIST_WARPS * 32);
    unsigned gx = (unsigned)(want < (size_t)sm_count * 8 ? (want ? want : 1) : (size_t)sm_count * 8);
    k_value_hist<<<gx, HIST_WARPS * 32, 0, s>>>(coeffs, n_coeffs, hist);
}

__global__ void __launch_bounds__(256)
k_heat_u8(const double* __restrict__ err, uint8_t* __restrict__ out, size_t n) {
    // four pixels per thread; the un-fused multiply and the clip are numpy's
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if (i4 + 4 <= n && ((uintptr_t)(out + i4) & 3) == 0) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double v = fmin(fmax(__dmul_rn(err[i4 + k], 10.0), 0.0), 255.0);
            w |= (uint32_t)(int)v << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(out + i4) = w;
    } else {
        for (size_t i = i4; i < n && i < i4 + 4; ++i)
            out[i] = (uint8_t)(int)fmin(fmax(__dmul_rn(err[i], 10.0), 0.0), 255.0);
    }
}

void launch_heat_u8(const double* err, uint8_t* out, size_t n, cudaStream_t s) {
    const size_t threads = (n + 3) / 4;
    k_heat_u8<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(err, out, n);
}

// ------------------------------------------------------------------------------
// stand-alone 8x8 block operators of engines/dct_engine.py:7-27 and
// engines/quantizer.py:22-29, exact arithmetic, one block per thread
// ------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_block_ops(int op, long long n_blocks, const double* __restrict__ in,
            const int16_t* __restrict__ in_q, const double* __restrict__ qtable,
            double* __restrict__ out, int16_t* __restrict__ out_q) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    typedef Exact P;
    double v[64];
    if (op == JDS_BLOCKOP_DEQUANTIZE) {
#pragma unroll
        for (int i = 0; i < 64; ++i) out[b * 64 + i] = P::mul((double)in_q[b * 64 + i], qtable[i]);
        return;
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = in[b * 64 + i];
    if (op == JDS_BLOCKOP_QUANTIZE) {
#pragma unroll
        for (int i = 0; i < 64; ++i) out_q[b * 64 + i] = (int16_t)(int)P::rint_(P::div(v[i], qtable[i]));
        return;
    }
    const bool forward = (op == JDS_BLOCKOP_DCT2 || op == JDS_BLOCKOP_ENCODE);
    if (op == JDS_BLOCKOP_ENCODE) {
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = P::sub(v[i], 128.0);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        double t[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) t[r] = v[r * 8 + c];
        if (forward) dct8_ref<P>(t, 0.0625); else idct8_ref<P>(t, 0.0625);
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r * 8 + c] = t[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (forward) dct8_ref<P>(v + r * 8, 1.0); else idct8_ref<P>(v + r * 8, 1.0);
    }
    if (op == JDS_BLOCKOP_DECODE) {
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = P::clamp255(P::add(v[i], 128.0));
    }
#pragma unroll
    for (int i = 0; i < 64; ++i) out[b * 64 + i] = v[i];
}

void launch_block_ops(int op, long long n_blocks, const double* in, const int16_t* in_q,
                      const double* qtable, double* out, int16_t* out_q, cudaStream_t s) {
    const unsigned grid = (unsigned)((n_blocks + 127) / 128);
    k_block_ops<<<grid, 128, 0, s>>>(op, n_blocks, in, in_q, qtable, out, out_q);
}

// ------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------
template <class P>
static void launch_forward_t(const Geom& g, int prefilter, const uint8_t* rgb, size_t rgb_stride,
                             typename P::T* fwd, size_t fwd_stride, int units, cudaStream_t s,
                             int chroma_only) {
    if (g.general) {
        dim3 blk(32, 8), grid((g.W + 31) / 32, (g.H + 7) / 8, units);
        if (prefilter)
            k_forward_general<P, true><<<grid, blk, 0, s>>>(g, rgb, rgb_stride, fwd, fwd_stride);
        else
            k_forward_general<P, false><<<grid, blk, 0, s>>>(g, rgb, rgb_stride, fwd, fwd_stride);
        return;
    }
    dim3 blk(32, 8), grid((g.wc + 31) / 32, (g.hc + 7) / 8, units);
    if (g.sub == 0)
        k_forward<P, 0, false><<<grid, blk, 0, s>>>(g, rgb, rgb_stride, fwd, fwd_stride, chroma_only);
    else if (g.sub == 1 && !prefilter)
        k_forward<P, 1, false><<<grid, blk, 0, s>>>(g, rgb, rgb_stride, fwd, fwd_stride, chroma_only);
    else if (g.sub == 1)
        k_forward<P, 1, true><<<grid, blk, 0, s>>>(g, rgb, rgb_stride, fwd, fwd_stride, chroma_only);
    else if (!prefilter)
        k_forward<P, 2, false><<<grid, blk, 0, s>>>(g, rgb, rgb_stride, fwd, fwd_stride, chroma_only);
    else
        k_forward<P, 2, true><<<grid, blk, 0, s>>>(g, rgb, rgb_stride, fwd, fwd_stride, chroma_only);
}

void launch_forward(bool exact, const Geom& g, int prefilter, const uint8_t* rgb,
                    size_t rgb_stride, void* fwd, size_t fwd_stride, int units, cudaStream_t s,
                    bool chroma_only) {
    // chroma_only (regular geometry only): the Y plane is neither computed nor written
    const int co = (chroma_only && !g.general) ? 1 : 0;
    if (exact)
        launch_forward_t<Exact>(g, prefilter, rgb, rgb_stride, (double*)fwd, fwd_stride, units, s, co);
    else
        launch_forward_t<Fast>(g, prefilter, rgb, rgb_stride, (float*)fwd, fwd_stride, units, s, co);
}

template <class P>
static void launch_codec_t(const Geom& g, const typename P::T* fwd, size_t fwd_stride,
                           typename P::T* rec, size_t rec_stride, const QTables* tables,
                           int table_stride, int16_t* coeffs, size_t coeff_stride, bool hist,
                           DevMetrics* metrics, int units, cudaStream_t s, bool chroma_only) {
    // chroma_only: the blocks of the Cb / Cr planes only (the fused exact luma kernel owns Y)
    const long long first = chroma_only ? g.nblk_y : 0;
    const long long total = g.nblk_y + 2 * g.nblk_c - first;
    dim3 blk(128), grid((unsigned)((total + 127) / 128), 1, units);
    if (hist)
        k_codec<P, true><<<grid, blk, 0, s>>>(g, fwd, fwd_stride, rec, rec_stride, tables,
                                              table_stride, coeffs, coeff_stride, metrics, first);
    else
        k_codec<P, false><<<grid, blk, 0, s>>>(g, fwd, fwd_stride, rec, rec_stride, tables,
                                               table_stride, coeffs, coeff_stride, metrics, first);
}

void launch_codec(bool exact, const Geom& g, const void* fwd, size_t fwd_stride, void* rec,
                  size_t rec_stride, const QTables* tables, int table_stride, int16_t* coeffs,
                  size_t coeff_stride, bool hist, DevMetrics* metrics, int units,
                  cudaStream_t s, bool chroma_only) {
    if (exact)
        launch_codec_t<Exact>(g, (const double*)fwd, fwd_stride, (double*)rec, rec_stride, tables,
                              table_stride, coeffs, coeff_stride, hist, metrics, units, s, chroma_only);
    else
        launch_codec_t<Fast>(g, (const float*)fwd, fwd_stride, (float*)rec, rec_stride, tables,
                             table_stride, coeffs, coeff_stride, hist, metrics, units, s, chroma_only);
}

void launch_inverse(bool exact, const Geom& g, const uint8_t* rgb, size_t rgb_stride,
                    const void* fwd, size_t fwd_stride, const void* rec, size_t rec_stride,
                    uint8_t* recon, size_t recon_stride, double* err_y, double* err_rgb,
                    DevMetrics* metrics, int units, cudaStream_t s) {
    const bool aligned = (g.W % 4) == 0 && (((uintptr_t)recon | recon_stride) & 3) == 0 &&
                         (((uintptr_t)err_y | (uintptr_t)err_rgb) & 15) == 0;
    if (aligned) {
        dim3 blk4(32, 8), grid4((g.W / 4 + 31) / 32, (g.H + 7) / 8, units);
        if (exact)
            k_inverse4<Exact><<<grid4, blk4, 0, s>>>(g, rgb, rgb_stride, (const double*)fwd, fwd_stride,
                                                     (const double*)rec, rec_stride, recon,
                                                     recon_stride, err_y, err_rgb, metrics);
        else
            k_inverse4<Fast><<<grid4, blk4, 0, s>>>(g, rgb, rgb_stride, (const float*)fwd, fwd_stride,
                                                    (const float*)rec, rec_stride, recon,
                                                    recon_stride, err_y, err_rgb, metrics);
        return;
    }
    dim3 blk(32, 8), grid((g.W + 31) / 32, (g.H + 7) / 8, units);
    if (exact)
        k_inverse<Exact><<<grid, blk, 0, s>>>(g, rgb, rgb_stride, (const double*)fwd, fwd_stride,
                                              (const double*)rec, rec_stride, recon, recon_stride,
                                              err_y, err_rgb, metrics);
    else
        k_inverse<Fast><<<grid, blk, 0, s>>>(g, rgb, rgb_stride, (const float*)fwd, fwd_stride,
                                             (const float*)rec, rec_stride, recon, recon_stride,
                                             err_y, err_rgb, metrics);
}

void launch_ssim(bool exact, int H, int W, const uint8_t* a, size_t a_stride, const uint8_t* b,
                 size_t b_stride, DevMetrics* metrics, int units, cudaStream_t s) {
    if (H < 7 || W < 7) return;
    dim3 blk(256), grid((W - 6 + SS_TW - 1) / SS_TW, (H - 6 + SS_TH - 1) / SS_TH, units);
    if (exact)
        k_ssim<double><<<grid, blk, 0, s>>>(H, W, a, a_stride, b, b_stride, metrics);
    else
        k_ssim<float><<<grid, blk, 0, s>>>(H, W, a, a_stride, b, b_stride, metrics);
}

void launch_selected_block(const Geom& g, const uint8_t* rgb, int bx, int by,
                           const QTables* tables, void* out, cudaStream_t s) {
    k_selected_block<<<1, 32, 0, s>>>(g, rgb, bx, by, tables, (SelectedOut*)out);
}

size_t selected_out_bytes() { return sizeof(SelectedOut); }

}  // namespace jds
