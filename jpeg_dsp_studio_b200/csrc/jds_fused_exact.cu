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

#ifndef JDS_XL_MIN_CTAS
#define JDS_XL_MIN_CTAS 3      // CTAs per SM the exact kernels are compiled for: 168 registers, 4-8 bytes of spills,
                               // 12 warps per SM (2: 255 registers, luma 0.96 ms; 3: 0.85 ms; 4: 1.01 ms per 16 x 4K)
#endif
#ifndef JDS_XL_COMPOSE_UNROLL
#define JDS_XL_COMPOSE_UNROLL 1
#endif
#ifndef JDS_XC_PREFETCH
#define JDS_XC_PREFETCH 1
#endif
#ifndef JDS_XL_PREFETCH
#define JDS_XL_PREFETCH 1
#endif
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

template <int SUB, bool COEFFS>
__global__ void __launch_bounds__(XL_NT, JDS_XL_MIN_CTAS)
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

#if JDS_XL_PREFETCH
    // The compose phase at the end reads this tile's reconstructed chroma straight from the planes
    // (16 frames of fp64 planes are not L2 resident): ask L2 for those lines now, so that the loads
    // find them there ~30 us later instead of waiting on DRAM one dependent task after the other.
    {
        constexpr int NCR = (SUB == 2) ? XL_TH / 2 + 2 : XL_TH;
        constexpr int CW = (SUB == 0) ? XL_TW : XL_TW / 2;         // chroma samples under the tile's row
        constexpr int LINES = CW * 8 / 128 + (SUB == 0 ? 0 : 1);   // 128-byte lines per row segment
        const int cr_lo = (SUB == 2) ? max((y0 >> 1) - 1, 0) : y0;
        const double* Cb0 = rec + (size_t)unit * rec_stride + g.plane_y;
        for (int i = tid; i < 2 * NCR * LINES; i += XL_NT) {
            const int chn = i / (NCR * LINES), r = (i / LINES) % NCR, l = i % LINES;
            const int row = min(cr_lo + r, g.hc - 1), col = (SUB == 0 ? x0 : (x0 >> 1)) + 16 * l;
            if (col < g.wcp) {
                const double* pl = Cb0 + (size_t)chn * g.plane_c + (size_t)row * g.wcp + col;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pl));
            }
        }
    }
#endif
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
    constexpr int COMPOSE_UNROLL = JDS_XL_COMPOSE_UNROLL;
#pragma unroll COMPOSE_UNROLL
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
            // horizontal taps of the four pixels x .. x+3 (x a multiple of 4, k = x / 2 even):
            //   x    : (S[k-1], S[k])   f = .75      x+1 : (S[k],   S[k+1]) f = .25
            //   x+2  : (S[k],   S[k+1]) f = .75      x+3 : (S[k+1], S[k+2]) f = .25
            // k-1 / k+2 clamp at the plane border (taps2) - four distinct samples per row, of which
            // S[k], S[k+1] are one aligned 16-byte load
            const int k = x >> 1;
            const int km1 = max(k - 1, 0), kp2 = min(k + 2, g.wc - 1);
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
                const double2 bm = __ldg(reinterpret_cast<const double2*>(pb + k));
                const double2 rm = __ldg(reinterpret_cast<const double2*>(pr + k));
                const double bl = __ldg(pb + km1), br = __ldg(pb + kp2);
                const double rl = __ldg(pr + km1), rr2 = __ldg(pr + kp2);
                const double db = P::sub(bm.y, bm.x), dr = P::sub(rm.y, rm.x);
                hb[rr][0] = P::fma(P::sub(bm.x, bl), 0.75, bl);
                hb[rr][1] = P::fma(db, 0.25, bm.x);
                hb[rr][2] = P::fma(db, 0.75, bm.x);
                hb[rr][3] = P::fma(P::sub(br, bm.y), 0.25, bm.y);
                hr[rr][0] = P::fma(P::sub(rm.x, rl), 0.75, rl);
                hr[rr][1] = P::fma(dr, 0.25, rm.x);
                hr[rr][2] = P::fma(dr, 0.75, rm.x);
                hr[rr][3] = P::fma(P::sub(rr2, rm.y), 0.25, rm.y);
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

// ------------------------------------------------------------------------------
// exact chroma kernel (no prefilter): RGB tile -> Cb / Cr (A1) -> 2x1 / 2x2 INTER_AREA
// average (A3) -> 8x8 codec -> reconstructed chroma planes, all in one pass; replaces the
// chroma-only k_forward + k_codec pair (two trips of the decimated planes through HBM and a
// codec whose threads each walk eight strided rows).  tile = 16 x 4 chroma blocks per channel;
// thread t: channel t / 64, block t % 64.
// ------------------------------------------------------------------------------
constexpr int XC_BX = 16, XC_BY = 4, XC_NT = 128;

struct ExactChromaSmem {
    alignas(16) double plane[2][XC_BX * XC_BY][XL_STRIDE];
    QTables tb;
};

// codec of the tile's 2 x 64 chroma blocks in place (thread t: channel t / 64, block t % 64) and
// the reconstructed planes out; shared by the plain and the prefilter chroma kernels
template <bool COEFFS>
__device__ __forceinline__ void exact_chroma_tail(const Geom& g, double (*plane)[XC_BX * XC_BY][XL_STRIDE],
                                                  const QTables& tb, int bx0, int by0, int unit,
                                                  double* __restrict__ rec, size_t rec_stride,
                                                  int16_t* __restrict__ coeffs, size_t coeff_stride,
                                                  DevMetrics* __restrict__ metrics) {
    typedef Exact P;
    const int tid = threadIdx.x;
    // ---- codec: thread t -> channel t / 64, block t % 64, in place ------------------------
    {
        const int ch = tid >> 6, blk = tid & 63;
        const int bx = bx0 + (blk & (XC_BX - 1)), by = by0 + (blk >> 4);
        unsigned long long bits = 0, nnz = 0;
        if (bx < g.nbx_c && by < g.nby_c) {
            double v[64];
            double2* slot = reinterpret_cast<double2*>(&plane[ch][blk][0]);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const double2 a = slot[i];
                v[2 * i] = a.x;
                v[2 * i + 1] = a.y;
            }
            int16_t q[64];
            BlockStats st;
            BlockCodec<P>::run(v, q, tb, st, nullptr, nullptr);
            bits = st.bits;
            nnz = st.nnz;
            if (COEFFS) {
                uint4* out = reinterpret_cast<uint4*>(
                    coeffs + (size_t)unit * coeff_stride +
                    ((size_t)g.nblk_y + (size_t)ch * g.nblk_c + (size_t)by * g.nbx_c + bx) * 64);
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
    // ---- reconstructed planes out: a task = (channel, chroma row, 8 samples): consecutive
    // threads write consecutive 64-byte pieces of a plane row --------------------------------
    for (int task = tid; task < 2 * XC_BY * 8 * XC_BX; task += XC_NT) {
        const int ch = task / (XC_BY * 8 * XC_BX);
        const int rem = task % (XC_BY * 8 * XC_BX);
        const int cr = rem / XC_BX, seg = rem % XC_BX;
        const int bx = bx0 + seg, cy = by0 * 8 + cr;
        if (bx >= g.nbx_c || cy >= g.hcp) continue;
        const int blk = (cr >> 3) * XC_BX + seg, ry = cr & 7;
        const double2* src = reinterpret_cast<const double2*>(&plane[ch][blk][ry * 8]);
        double2* dst = reinterpret_cast<double2*>(rec + (size_t)unit * rec_stride + g.plane_y +
                                                  (size_t)ch * g.plane_c + (size_t)cy * g.wcp + bx * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = src[i];
    }
}


template <int SUB, bool COEFFS>
__global__ void __launch_bounds__(XC_NT, JDS_XL_MIN_CTAS)
k_exact_chroma(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
               double* __restrict__ rec, size_t rec_stride,
               const QTables* __restrict__ tables, int table_stride,
               int16_t* __restrict__ coeffs, size_t coeff_stride, DevMetrics* __restrict__ metrics) {
    typedef Exact P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ExactChromaSmem& sm = *reinterpret_cast<ExactChromaSmem*>(smem_raw);
    constexpr int VS = (SUB == 2) ? 2 : 1;            // luma rows per chroma row
    constexpr int HS = (SUB == 0) ? 1 : 2;            // luma pixels per chroma sample (4:4:4: the planes are the pixels)
    constexpr int SEG_PX = 8 * HS;                    // luma pixels of a task (8 chroma samples)
    const int tid = threadIdx.x;
    const int unit = blockIdx.z;
    const uint8_t* in = rgb + (size_t)unit * rgb_stride;
    const int bx0 = blockIdx.x * XC_BX, by0 = blockIdx.y * XC_BY;   // chroma block origin
    const int x0 = bx0 * SEG_PX, y0 = by0 * 8 * VS;                  // luma pixel origin
    const int n_rows = min(XC_BY * 8 * VS, g.H - y0);
    const int n_px = min(XC_BX * SEG_PX, g.W - x0);
    {
        const QTables* src = tables + (size_t)unit * table_stride;
        for (int i = tid; i < 64; i += XC_NT) {
            sm.tb.q[i] = src->q[i];
            sm.tb.rq[i] = src->rq[i];
            sm.tb.dqx[i] = src->dqx[i];
        }
    }
#if JDS_XC_PREFETCH
    // the four tasks below load their rows one after the other: have L2 fetch the whole tile now
    {
        constexpr int LINES = XC_BX * SEG_PX * 3 / 128;             // 6 (3) lines of 128 B per luma row
        const size_t row_end = (size_t)g.W * 3;
        for (int i = tid; i < XC_BY * 8 * VS * LINES; i += XC_NT) {
            const int r = i / LINES, l = i % LINES;
            const size_t off = (size_t)x0 * 3 + 128 * (size_t)l;
            if (r < n_rows && off < row_end)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(in + (size_t)(y0 + r) * row_end + off));
        }
    }
#endif
    // ---- decimation: a task = one chroma row x 8 chroma samples (16 / 8 luma pixels) -----
    {
        constexpr int NTASK = XC_BY * 8 * XC_BX / XC_NT;            // 4
        const int seg = tid & 15, crbase = tid >> 4;
        const bool seg_ok = seg * SEG_PX < n_px;
#pragma unroll 1
        for (int k = 0; k < NTASK; ++k) {
            const int cr = crbase + 8 * k;
            if (!(seg_ok && cr * VS < n_rows)) continue;
            uint32_t w[VS][12];
#pragma unroll
            for (int v = 0; v < VS; ++v) {
                const uint8_t* row = in + ((size_t)(y0 + cr * VS + v) * g.W + x0 + seg * SEG_PX) * 3;
                if (SUB == 0) {                         // 24 bytes, 8-byte aligned
                    const uint2* q = reinterpret_cast<const uint2*>(row);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const uint2 a = __ldg(q + i);
                        w[v][2 * i] = a.x; w[v][2 * i + 1] = a.y;
                    }
                } else {
                    const uint4* q = reinterpret_cast<const uint4*>(row);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const uint4 a = __ldg(q + i);
                        w[v][4 * i] = a.x; w[v][4 * i + 1] = a.y; w[v][4 * i + 2] = a.z; w[v][4 * i + 3] = a.w;
                    }
                }
            }
            const int blk = (cr >> 3) * XC_BX + seg, ry = cr & 7;
            double2* pb = reinterpret_cast<double2*>(&sm.plane[0][blk][ry * 8]);
            double2* pr = reinterpret_cast<double2*>(&sm.plane[1][blk][ry * 8]);
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {            // two chroma samples per 16-byte store
                double ob[2], orr[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int px = HS * (2 * s2 + e);   // first luma pixel of the sample
                    double cb[VS][2], cr_[VS][2];
#pragma unroll
                    for (int v = 0; v < VS; ++v)
#pragma unroll
                        for (int d = 0; d < HS; ++d) {
                            const int b = 3 * (px + d);
                            const double rr = byte_to_double(w[v][b >> 2], b & 3);
                            const double gg = byte_to_double(w[v][(b + 1) >> 2], (b + 1) & 3);
                            const double bb = byte_to_double(w[v][(b + 2) >> 2], (b + 2) & 3);
                            rgb_to_cbcr<P>(rr, gg, bb, cb[v][d], cr_[v][d]);
                        }
                    if (SUB == 0) {                     // 4:4:4: subsample_chroma copies (color_space.py:35-36)
                        ob[e] = cb[0][0];
                        orr[e] = cr_[0][0];
                    } else if (SUB == 1) {              // A3, 4:2:2: (a+b)*0.5
                        ob[e] = P::mul(P::add(cb[0][0], cb[0][1]), 0.5);
                        orr[e] = P::mul(P::add(cr_[0][0], cr_[0][1]), 0.5);
                    } else {                            // A3, 4:2:0: (((a+b)+c)+d)*0.25
                        ob[e] = P::mul(P::add(P::add(P::add(cb[0][0], cb[0][1]), cb[VS - 1][0]), cb[VS - 1][1]), 0.25);
                        orr[e] = P::mul(P::add(P::add(P::add(cr_[0][0], cr_[0][1]), cr_[VS - 1][0]), cr_[VS - 1][1]), 0.25);
                    }
                }
                pb[s2] = make_double2(ob[0], ob[1]);
                pr[s2] = make_double2(orr[0], orr[1]);
            }
        }
    }
    __syncthreads();
    exact_chroma_tail<COEFFS>(g, sm.plane, sm.tb, bx0, by0, unit, rec, rec_stride, coeffs, coeff_stride, metrics);
}

// ------------------------------------------------------------------------------
// exact chroma kernel WITH the anti-alias prefilter (use_prefilter=True, 4:2:2 / 4:2:0):
// cv2.GaussianBlur(3x3, sigma 0.75, REFLECT_101) of the full-resolution Cb / Cr (A2) and the
// INTER_AREA average (A3) in the reference's own operation order - the blur cannot be folded
// into the average here, every product and sum is individually rounded.  The tile's luma rows
// go through the row pass in passes of XP_PROWS rows (+1 halo row above and below) into a
// shared-memory buffer; the column pass and the average then produce the pass's chroma rows.
// Replaces k_forward(chroma_only) - which recomputes the colour conversion and the row pass of
// the 4 x 4 pixels under every chroma sample - plus the chroma-only k_codec.
// Block-aligned frames only (fused_supported: W % 16 == 0, so A2's non-FMA tail columns do
// not occur: W4 == W).
// ------------------------------------------------------------------------------
constexpr int XP_PROWS = 8;                               // luma rows per pass

template <int SUB>
struct ExactChromaPfSmem {
    static constexpr int VS = (SUB == 2) ? 2 : 1;
    static constexpr int LROWS = XC_BY * 8 * VS;          // luma rows of the tile
    static constexpr int PARTS = LROWS / XP_PROWS;
    alignas(16) double plane[2][XC_BX * XC_BY][XL_STRIDE];
    QTables tb;
    // row-filtered Cb / Cr of the pass's luma rows (row 0 = the row above the pass), pixel pair i of
    // segment seg at [i][seg]: consecutive lanes (segments) touch consecutive 16-byte slots
    alignas(16) double2 hrow[XP_PROWS + 2][2][8][XC_BX];
};

template <int SUB, bool COEFFS>
__global__ void __launch_bounds__(XC_NT, 2)
k_exact_chroma_pf(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
                  double* __restrict__ rec, size_t rec_stride,
                  const QTables* __restrict__ tables, int table_stride,
                  int16_t* __restrict__ coeffs, size_t coeff_stride, DevMetrics* __restrict__ metrics) {
    typedef Exact P;
    typedef ExactChromaPfSmem<SUB> Smem;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    constexpr int VS = Smem::VS;
    const int tid = threadIdx.x;
    const int unit = blockIdx.z;
    const uint8_t* in = rgb + (size_t)unit * rgb_stride;
    const int bx0 = blockIdx.x * XC_BX, by0 = blockIdx.y * XC_BY;   // chroma block origin
    const int x0 = bx0 * 16, y0 = by0 * 8 * VS;                      // luma pixel origin
    const int n_rows = min(Smem::LROWS, g.H - y0);
    const int n_px = min(XC_BX * 16, g.W - x0);
    {
        const QTables* src = tables + (size_t)unit * table_stride;
        for (int i = tid; i < 64; i += XC_NT) {
            sm.tb.q[i] = src->q[i];
            sm.tb.rq[i] = src->rq[i];
            sm.tb.dqx[i] = src->dqx[i];
        }
    }
#if JDS_XC_PREFETCH
    {   // rows of the later passes: ask L2 for them while the first pass runs
        constexpr int LINES = XC_BX * 16 * 3 / 128;
        const size_t row_end = (size_t)g.W * 3;
        for (int i = tid; i < Smem::LROWS * LINES; i += XC_NT) {
            const int r = i / LINES, l = i % LINES;
            const size_t off = (size_t)x0 * 3 + 128 * (size_t)l;
            if (r >= XP_PROWS && r < n_rows && off < row_end)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(in + (size_t)(y0 + r) * row_end + off));
        }
    }
#endif
    for (int part = 0; part < Smem::PARTS; ++part) {
        const int lr0 = part * XP_PROWS;                  // first luma row of the pass (tile local)
        if (lr0 >= n_rows) break;
        if (part) __syncthreads();                        // the last pass's phase B is done with hrow
        // ---- phase A: a task = one luma row (halo rows included) x 16 pixels: colour (A1) and
        // the row pass of the blur (A2) ------------------------------------------------------
        for (int task = tid; task < (XP_PROWS + 2) * XC_BX; task += XC_NT) {
            const int hr = task / XC_BX, seg = task % XC_BX;
            if (lr0 + hr - 1 >= n_rows + 1 || seg * 16 >= n_px) continue;
            const int y = reflect101(y0 + lr0 + hr - 1, g.H);          // REFLECT_101 at the frame's top / bottom
            const int xs = x0 + seg * 16;
            const uint8_t* row = in + ((size_t)y * g.W + xs) * 3;
            const bool has_l = xs > 0, has_r = xs + 16 < g.W;
            const uint4* q = reinterpret_cast<const uint4*>(row);
            const uint4 c0 = __ldg(q), c1 = __ldg(q + 1), c2 = __ldg(q + 2);
            const uint32_t wl = has_l ? __ldg(reinterpret_cast<const uint32_t*>(row) - 1) : 0u;   // bytes -4..-1
            const uint32_t wr = has_r ? __ldg(reinterpret_cast<const uint32_t*>(row) + 12) : 0u;  // bytes 48..51
            const uint32_t w[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
            double cb[18], cr[18];                        // index i <-> pixel xs - 1 + i
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int b = 3 * i;
                rgb_to_cbcr<P>(byte_to_double(w[b >> 2], b & 3), byte_to_double(w[(b + 1) >> 2], (b + 1) & 3),
                               byte_to_double(w[(b + 2) >> 2], (b + 2) & 3), cb[i + 1], cr[i + 1]);
            }
            if (has_l) rgb_to_cbcr<P>(byte_to_double(wl, 1), byte_to_double(wl, 2), byte_to_double(wl, 3), cb[0], cr[0]);
            else { cb[0] = cb[2]; cr[0] = cr[2]; }        // REFLECT_101: pixel -1 = pixel 1
            if (has_r) rgb_to_cbcr<P>(byte_to_double(wr, 0), byte_to_double(wr, 1), byte_to_double(wr, 2), cb[17], cr[17]);
            else { cb[17] = cb[15]; cr[17] = cr[15]; }    // pixel W = pixel W - 2
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                sm.hrow[hr][0][i >> 1][seg] = make_double2(blur_row<P>(cb[i], cb[i + 1], cb[i + 2], false),
                                                           blur_row<P>(cb[i + 1], cb[i + 2], cb[i + 3], false));
                sm.hrow[hr][1][i >> 1][seg] = make_double2(blur_row<P>(cr[i], cr[i + 1], cr[i + 2], false),
                                                           blur_row<P>(cr[i + 1], cr[i + 2], cr[i + 3], false));
            }
        }
        __syncthreads();
        // ---- phase B: a task = (channel, chroma row of the pass, 8 chroma samples): column pass
        // (A2) on the VS luma rows under the samples, then the INTER_AREA average (A3) -------------
        for (int task = tid; task < 2 * (XP_PROWS / VS) * XC_BX; task += XC_NT) {
            const int ch = task / ((XP_PROWS / VS) * XC_BX);
            const int rem = task % ((XP_PROWS / VS) * XC_BX);
            const int crp = rem / XC_BX, seg = rem % XC_BX;               // chroma row inside the pass
            const int lr = lr0 + crp * VS;                                // its first luma row (tile local)
            if (lr >= n_rows || seg * 16 >= n_px) continue;
            double f[VS][16];
#pragma unroll
            for (int v = 0; v < VS; ++v) {
                const int hr = crp * VS + v + 1;                          // buffer row of luma row lr + v
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const double2 a = sm.hrow[hr - 1][ch][i][seg], b = sm.hrow[hr][ch][i][seg],
                                  c = sm.hrow[hr + 1][ch][i][seg];
                    f[v][2 * i] = blur_col<P>(a.x, b.x, c.x);
                    f[v][2 * i + 1] = blur_col<P>(a.y, b.y, c.y);
                }
            }
            const int cr_t = (lr0 / VS) + crp;                            // chroma row inside the tile
            const int blk = (cr_t >> 3) * XC_BX + seg, ry = cr_t & 7;
            double2* dst = reinterpret_cast<double2*>(&sm.plane[ch][blk][ry * 8]);
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {
                double o[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int px = 2 * (2 * s2 + e);
                    if (SUB == 1) o[e] = P::mul(P::add(f[0][px], f[0][px + 1]), 0.5);
                    else o[e] = P::mul(P::add(P::add(P::add(f[0][px], f[0][px + 1]), f[VS - 1][px]), f[VS - 1][px + 1]), 0.25);
                }
                dst[s2] = make_double2(o[0], o[1]);
            }
        }
    }
    __syncthreads();
    exact_chroma_tail<COEFFS>(g, sm.plane, sm.tb, bx0, by0, unit, rec, rec_stride, coeffs, coeff_stride, metrics);
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
#define JDS_SET(K)                                                                              \
    if ((e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize,                \
                                  (int)sizeof(ExactChromaSmem))) != cudaSuccess) return e
    JDS_SET((k_exact_chroma<0, false>)); JDS_SET((k_exact_chroma<0, true>));
    JDS_SET((k_exact_chroma<1, false>)); JDS_SET((k_exact_chroma<1, true>));
    JDS_SET((k_exact_chroma<2, false>)); JDS_SET((k_exact_chroma<2, true>));
#undef JDS_SET
#define JDS_SET(K, BYTES)                                                                       \
    if ((e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES))) != cudaSuccess) return e
    JDS_SET((k_exact_chroma_pf<1, false>), sizeof(ExactChromaPfSmem<1>));
    JDS_SET((k_exact_chroma_pf<1, true>), sizeof(ExactChromaPfSmem<1>));
    JDS_SET((k_exact_chroma_pf<2, false>), sizeof(ExactChromaPfSmem<2>));
    JDS_SET((k_exact_chroma_pf<2, true>), sizeof(ExactChromaPfSmem<2>));
#undef JDS_SET
    return cudaSuccess;
}

// chroma of a block-aligned frame in one kernel (4:4:4 has no decimation, hence no prefilter)
bool exact_chroma_supported(const Geom& g, int prefilter) {
    (void)prefilter;
    return g.sub == 0 || !g.general;
}

cudaError_t launch_exact_chroma(const Geom& g, int prefilter, const uint8_t* rgb, size_t rgb_stride, double* rec,
                                size_t rec_stride, const QTables* tables, int table_stride,
                                int16_t* coeffs, size_t coeff_stride, DevMetrics* metrics, int units,
                                cudaStream_t s) {
    dim3 grid((g.nbx_c + XC_BX - 1) / XC_BX, (g.nby_c + XC_BY - 1) / XC_BY, units);
    if (prefilter && g.sub != 0) {
#define JDS_LAUNCH_XP(SUBV, CO)                                                                 \
    k_exact_chroma_pf<SUBV, CO><<<grid, XC_NT, sizeof(ExactChromaPfSmem<SUBV>), s>>>(            \
        g, rgb, rgb_stride, rec, rec_stride, tables, table_stride, coeffs, coeff_stride, metrics)
        if (g.sub == 1) { if (coeffs) JDS_LAUNCH_XP(1, true); else JDS_LAUNCH_XP(1, false); }
        else { if (coeffs) JDS_LAUNCH_XP(2, true); else JDS_LAUNCH_XP(2, false); }
#undef JDS_LAUNCH_XP
        return cudaGetLastError();
    }
    const size_t smem = sizeof(ExactChromaSmem);
#define JDS_LAUNCH_XC(SUBV, CO)                                                                 \
    k_exact_chroma<SUBV, CO><<<grid, XC_NT, smem, s>>>(g, rgb, rgb_stride, rec, rec_stride, tables, \
                                                       table_stride, coeffs, coeff_stride, metrics)
    if (g.sub == 0) { if (coeffs) JDS_LAUNCH_XC(0, true); else JDS_LAUNCH_XC(0, false); }
    else if (g.sub == 1) { if (coeffs) JDS_LAUNCH_XC(1, true); else JDS_LAUNCH_XC(1, false); }
    else { if (coeffs) JDS_LAUNCH_XC(2, true); else JDS_LAUNCH_XC(2, false); }
#undef JDS_LAUNCH_XC
    return cudaGetLastError();
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
