// jds_fused_exact.cu - fused EXACT-mode (fp64, reference arithmetic) luma + compose kernel.
//
// Exact mode reproduces engines/pipeline.py:47-95 bit for bit; the staged
// kernels of jds_kernels.cu do that through full-resolution fp64 planes in HBM (Y forward, Y
// reconstructed: 16 B/px written and 24 B/px read on top of the images).  On block-aligned
// frames this kernel removes the luma planes altogether:
//
//   k_forward (chroma only)   RGB -> decimated Cb / Cr planes            (jds_kernels.cu)
//   k_codec   (chroma only)   8x8 codec of the chroma blocks -> rec planes
//   k_exact_luma<SUB>         RGB tile -> Y (A1) -> 8x8 codec in registers (A5-A7: the very
//                             BlockCodec<Exact> the staged kernel runs) -> chroma upsample from
//                             the reconstructed planes (A8) -> YCbCr->RGB, clamp, truncate (A9)
//                             -> packed uint8 RGB
//
// Squared errors and SSIM come from k_ssim_strip (integer SSE: psnr_rgb stays bit-identical).
// The arithmetic is the policy code of jds_math.cuh / jds_stages.cuh - every operation
// individually rounded, in the reference's order - so coefficients and pixels equal the staged
// path's (and the reference's) bit for bit; only the data movement differs.
#include <cuda_runtime.h>
#include <stdint.h>
#include "jds_kernels.cuh"
#include "jds_stages.cuh"

namespace jds {

constexpr int XL_BX = 32, XL_BY = 4, XL_NT = 128;     // 32 x 4 luma blocks = 256 x 32 pixels per CTA
constexpr int XL_TW = XL_BX * 8, XL_TH = XL_BY * 8;
constexpr int XL_STRIDE = 66;                         // doubles per block slot: 528 B = 4 * 128 + 16,
                                                      // consecutive slots sit 16 B apart modulo 128

struct ExactLumaSmem {
    alignas(16) double plane[XL_BX * XL_BY][XL_STRIDE];
    QTables tb;
};

// byte `k` of `w` as fp64 without the conversion unit: 2^52 + v assembled from its bit
// pattern (one PRMT), minus 2^52 (one exact DADD)
__device__ __forceinline__ double byte_to_double(uint32_t w, int k) {
    return __dsub_rn(__hiloint2double(0x43300000, (int)__byte_perm(w, 0u, 0x4440 | k)), 4503599627370496.0);
}
// floor of a value in [0, 255] as an integer: 2^52 + x rounded DOWN keeps floor(x) in the low word
__device__ __forceinline__ uint32_t trunc_u8(double x) {
    return (uint32_t)__double2loint(__dadd_rd(x, 4503599627370496.0));
}

// cv2.resize(INTER_LINEAR) source taps at the exact factor 2 (A8): destination index i reads
// samples i0, i1 with weight f on (S[i1] - S[i0]); f = fma(i + .5, .5, -.5) - floor(.) is
// exactly .75 for even and .25 for odd i, the indices clamp at the plane border
__device__ __forceinline__ void taps2(int i, int n, int& i0, int& i1, double& f) {
    const int k = i >> 1;
    if (i & 1) {
        i0 = k;
        i1 = min(k + 1, n - 1);
        f = 0.25;
    } else {
        i0 = max(k - 1, 0);
        i1 = k;
        f = 0.75;
    }
}

template <int SUB, bool COEFFS>
__global__ void __launch_bounds__(XL_NT)
k_exact_luma(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
             const double* __restrict__ rec, size_t rec_stride,
             const QTables* __restrict__ tables, int table_stride,
             int16_t* __restrict__ coeffs, size_t coeff_stride,
             uint8_t* __restrict__ recon, size_t recon_stride, DevMetrics* __restrict__ metrics) {
    typedef Exact P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ExactLumaSmem& sm = *reinterpret_cast<ExactLumaSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int unit = blockIdx.z;
    const uint8_t* in = rgb + (size_t)unit * rgb_stride;
    const int x0 = blockIdx.x * XL_TW, y0 = blockIdx.y * XL_TH;
    const int n_rows = min(XL_TH, g.H - y0);
    const int n_px = min(XL_TW, g.W - x0);
    {
        const QTables* src = tables + (size_t)unit * table_stride;
        for (int i = tid; i < 64; i += XL_NT) {
            sm.tb.q[i] = src->q[i];
            sm.tb.rq[i] = src->rq[i];
            sm.tb.dqx[i] = src->dqx[i];
        }
    }

    // ---- RGB -> Y (A1), block layout: a task = one row x 16 pixels ---------------------
    {
        constexpr int NTASK = XL_TH * (XL_TW / 16) / XL_NT;        // 4
        const int seg = tid & 15, rbase = tid >> 4;
        const bool seg_ok = seg * 16 < n_px;
        uint4 ld[NTASK][3];
#pragma unroll
        for (int k = 0; k < NTASK; ++k) {
            const int r = rbase + 8 * k;
            if (seg_ok && r < n_rows) {
                const uint4* q = reinterpret_cast<const uint4*>(in + ((size_t)(y0 + r) * g.W + x0 + seg * 16) * 3);
                ld[k][0] = __ldg(q);
                ld[k][1] = __ldg(q + 1);
                ld[k][2] = __ldg(q + 2);
            }
        }
#pragma unroll
        for (int k = 0; k < NTASK; ++k) {
            const int r = rbase + 8 * k;
            if (seg_ok && r < n_rows) {
                const uint32_t w[12] = {ld[k][0].x, ld[k][0].y, ld[k][0].z, ld[k][0].w,
                                        ld[k][1].x, ld[k][1].y, ld[k][1].z, ld[k][1].w,
                                        ld[k][2].x, ld[k][2].y, ld[k][2].z, ld[k][2].w};
                const int blk = (r >> 3) * XL_BX + seg * 2, ry = r & 7;
                double2* p0 = reinterpret_cast<double2*>(&sm.plane[blk][ry * 8]);
                double2* p1 = reinterpret_cast<double2*>(&sm.plane[blk + 1][ry * 8]);
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    double yv[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int b = 3 * (i + e);
                        const double rr = byte_to_double(w[b >> 2], b & 3);
                        const double gg = byte_to_double(w[(b + 1) >> 2], (b + 1) & 3);
                        const double bb = byte_to_double(w[(b + 2) >> 2], (b + 2) & 3);
                        yv[e] = luma601<P>(rr, gg, bb);
                    }
                    if (i < 8) p0[i >> 1] = make_double2(yv[0], yv[1]);
                    else p1[(i - 8) >> 1] = make_double2(yv[0], yv[1]);
                }
            }
        }
    }
    __syncthreads();

    // ---- codec: one luma block per thread, in place (the staged kernel's BlockCodec) -----
    {
        const int bx = (x0 >> 3) + (tid & (XL_BX - 1)), by = (y0 >> 3) + (tid >> 5);
        unsigned long long bits = 0, nnz = 0;
        if (bx < g.nbx_y && by < g.nby_y) {
            double v[64];
            double2* slot = reinterpret_cast<double2*>(&sm.plane[tid][0]);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const double2 a = slot[i];
                v[2 * i] = a.x;
                v[2 * i + 1] = a.y;
            }
            int16_t q[64];
            BlockStats st;
            BlockCodec<P>::run(v, q, sm.tb, st, nullptr, nullptr);
            bits = st.bits;
            nnz = st.nnz;
            if (COEFFS) {
                uint4* out = reinterpret_cast<uint4*>(coeffs + (size_t)unit * coeff_stride +
                                                      ((size_t)by * g.nbx_y + bx) * 64);
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
#pragma unroll
            for (int i = 0; i < 32; ++i) slot[i] = make_double2(v[2 * i], v[2 * i + 1]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            bits += __shfl_down_sync(0xffffffffu, bits, o);
            nnz += __shfl_down_sync(0xffffffffu, nnz, o);
        }
        if ((tid & 31) == 0 && nnz) {
            atomicAdd(&metrics[unit].coeff_bits, bits);
            atomicAdd(&metrics[unit].nnz, nnz);
        }
    }
    __syncthreads();

    // ---- compose: a task = RPT rows x 4 pixels --------------------------------------------
    // chroma comes straight from the reconstructed planes (written by k_codec earlier in this
    // stream, L2 resident): every sample is read by the 2 x 2 pixels under it and their
    // neighbours, so the loads hit L1 after the first touch
    constexpr int RPT = (SUB == 2) ? 2 : 1;
    const double* Cbr = rec + (size_t)unit * rec_stride + g.plane_y;
    const double* Crr = Cbr + g.plane_c;
    uint8_t* out = recon + (size_t)unit * recon_stride;
    for (int task = tid; task < (XL_TH / RPT) * (XL_TW / 4); task += XL_NT) {
        const int rp = task / (XL_TW / 4), g4 = task % (XL_TW / 4);
        const int r0 = rp * RPT;
        if (r0 >= n_rows || g4 * 4 >= n_px) continue;
        const int x = x0 + 4 * g4;
        double cb[RPT][4], cr[RPT][4];
        if (SUB == 0) {
            const double* pb = Cbr + (size_t)(y0 + r0) * g.wcp + x;
            const double* pr = Crr + (size_t)(y0 + r0) * g.wcp + x;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                cb[0][i] = __ldg(pb + i);
                cr[0][i] = __ldg(pr + i);
            }
        } else {
            // horizontal taps of the four pixels (x is a multiple of 4)
            int i0[4], i1[4];
            double fx[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) taps2(x + i, g.wc, i0[i], i1[i], fx[i]);
            // chroma rows: 4:2:0 - the three rows around the pair (y0r, y1r per luma row);
            // 4:2:2 - the luma row itself
            constexpr int NCR = (SUB == 2) ? 3 : 1;
            int crow[NCR];
            if (SUB == 2) {
                const int j = (y0 + r0) >> 1;
                crow[0] = max(j - 1, 0);
                crow[1] = j;
                crow[2] = min(j + 1, g.hc - 1);
            } else {
                crow[0] = y0 + r0;
            }
            double hb[NCR][4], hr[NCR][4];
#pragma unroll
            for (int rr = 0; rr < NCR; ++rr) {
                const double* pb = Cbr + (size_t)crow[rr] * g.wcp;
                const double* pr = Crr + (size_t)crow[rr] * g.wcp;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double b0 = __ldg(pb + i0[i]), b1 = __ldg(pb + i1[i]);
                    const double c0 = __ldg(pr + i0[i]), c1 = __ldg(pr + i1[i]);
                    hb[rr][i] = P::fma(P::sub(b1, b0), fx[i], b0);
                    hr[rr][i] = P::fma(P::sub(c1, c0), fx[i], c0);
                }
            }
            if (SUB == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // even luma row: rows (j-1, j), weight .75; odd: rows (j, j+1), weight .25
                    cb[0][i] = P::fma(P::sub(hb[1][i], hb[0][i]), 0.75, hb[0][i]);
                    cr[0][i] = P::fma(P::sub(hr[1][i], hr[0][i]), 0.75, hr[0][i]);
                    cb[RPT - 1][i] = P::fma(P::sub(hb[NCR - 1][i], hb[1 % NCR][i]), 0.25, hb[1 % NCR][i]);
                    cr[RPT - 1][i] = P::fma(P::sub(hr[NCR - 1][i], hr[1 % NCR][i]), 0.25, hr[1 % NCR][i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    cb[0][i] = hb[0][i];
                    cr[0][i] = hr[0][i];
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
            const int r = r0 + rr;
            if (r >= n_rows) break;
            const int blk = (r >> 3) * XL_BX + (g4 >> 1), ry = r & 7;
            const double2* py = reinterpret_cast<const double2*>(&sm.plane[blk][ry * 8 + 4 * (g4 & 1)]);
            const double2 ya = py[0], yb = py[1];
            const double yv[4] = {ya.x, ya.y, yb.x, yb.y};
            uint32_t by[12];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double rf, gf, bf;
                ycbcr_to_rgb<P>(yv[i], cb[rr][i], cr[rr][i], rf, gf, bf);
                by[3 * i] = trunc_u8(rf);                 // .astype(uint8): truncation (A9)
                by[3 * i + 1] = trunc_u8(gf);
                by[3 * i + 2] = trunc_u8(bf);
            }
            uint32_t* dst = reinterpret_cast<uint32_t*>(out + ((size_t)(y0 + r) * g.W + x) * 3);
#pragma unroll
            for (int i = 0; i < 3; ++i)
                dst[i] = by[4 * i] | (by[4 * i + 1] << 8) | (by[4 * i + 2] << 16) | (by[4 * i + 3] << 24);
        }
    }
}

cudaError_t exact_fused_configure_device() {
    cudaError_t e;
#define JDS_SET(K)                                                                              \
    if ((e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize,                \
                                  (int)sizeof(ExactLumaSmem))) != cudaSuccess) return e
    JDS_SET((k_exact_luma<0, false>)); JDS_SET((k_exact_luma<0, true>));
    JDS_SET((k_exact_luma<1, false>)); JDS_SET((k_exact_luma<1, true>));
    JDS_SET((k_exact_luma<2, false>)); JDS_SET((k_exact_luma<2, true>));
#undef JDS_SET
    return cudaSuccess;
}

cudaError_t launch_exact_luma(const Geom& g, const uint8_t* rgb, size_t rgb_stride, const double* rec,
                              size_t rec_stride, const QTables* tables, int table_stride,
                              int16_t* coeffs, size_t coeff_stride, uint8_t* recon,
                              size_t recon_stride, DevMetrics* metrics, int units, cudaStream_t s) {
    dim3 grid((g.nbx_y + XL_BX - 1) / XL_BX, (g.nby_y + XL_BY - 1) / XL_BY, units);
    const size_t smem = sizeof(ExactLumaSmem);
#define JDS_LAUNCH_XL(SUBV, CO)                                                                 \
    k_exact_luma<SUBV, CO><<<grid, XL_NT, smem, s>>>(g, rgb, rgb_stride, rec, rec_stride, tables, \
                                                     table_stride, coeffs, coeff_stride, recon,  \
                                                     recon_stride, metrics)
    if (g.sub == 0) { if (coeffs) JDS_LAUNCH_XL(0, true); else JDS_LAUNCH_XL(0, false); }
    else if (g.sub == 1) { if (coeffs) JDS_LAUNCH_XL(1, true); else JDS_LAUNCH_XL(1, false); }
    else { if (coeffs) JDS_LAUNCH_XL(2, true); else JDS_LAUNCH_XL(2, false); }
#undef JDS_LAUNCH_XL
    return cudaGetLastError();
}

}  // namespace jds
