// emul.cpp - TEST INFRASTRUCTURE: runs the product's per-work-item device functions
// (jpeg_dsp_studio_b200/csrc/jds_stages.cuh, the bodies of the CUDA kernels) on the CPU,
// one call per thread of the corresponding kernel, so that the arithmetic can be
// checked against the oracle in the GPU-less build container.  Never linked into
// libjds.so and never imported by the product package.
//
// Build (tests/emul/build.py): g++ -O2 -ffp-contract=off (no -mfma / -march=native),
// so every fp64 operation of the Exact policy is individually rounded as on the GPU.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../jpeg_dsp_studio_b200/csrc/jds_host.h"
#include "../../jpeg_dsp_studio_b200/csrc/jds_stages.cuh"

using namespace jds;

template <class P, int SUB, bool PF>
static void run_forward(const Geom& g, const uint8_t* rgb, typename P::T* Y, typename P::T* Cb,
                        typename P::T* Cr) {
    for (int cy = 0; cy < g.hc; ++cy)
        for (int cx = 0; cx < g.wc; ++cx) forward_cell<P, SUB, PF>(g, rgb, cx, cy, Y, Cb, Cr);
}

template <class P>
static int run(int H, int W, int quality, int sub, int prefilter, const uint8_t* rgb,
               uint8_t* recon, int16_t* coeffs, double* err_y, double* err_rgb, uint64_t* bits,
               uint64_t* nnz, int64_t* hist, uint64_t* sse_rgb, double* sse_y) {
    typedef typename P::T T;
    Geom g;
    int rc = geom_init(H, W, sub, &g);
    if (rc) return -rc;
    QTables tb;
    fill_tables(quality, &tb);
    const size_t pe = (size_t)(g.plane_y + 2 * g.plane_c);
    std::vector<T> fwd(pe, T(0)), rec(pe, T(0));
    T *Y = fwd.data(), *Cb = Y + g.plane_y, *Cr = Cb + g.plane_c;
    if (g.general) {
        for (int y = 0; y < g.H; ++y)
            for (int x = 0; x < g.W; ++x) {
                forward_luma<P>(g, rgb, x, y, Y);
                if (x < g.wc && y < g.hc) {
                    if (prefilter) forward_chroma_area<P, true>(g, rgb, x, y, Cb, Cr);
                    else forward_chroma_area<P, false>(g, rgb, x, y, Cb, Cr);
                }
            }
    }
    else if (sub == 0) run_forward<P, 0, false>(g, rgb, Y, Cb, Cr);
    else if (sub == 1 && !prefilter) run_forward<P, 1, false>(g, rgb, Y, Cb, Cr);
    else if (sub == 1) run_forward<P, 1, true>(g, rgb, Y, Cb, Cr);
    else if (!prefilter) run_forward<P, 2, false>(g, rgb, Y, Cb, Cr);
    else run_forward<P, 2, true>(g, rgb, Y, Cb, Cr);

    *bits = 0;
    *nnz = 0;
    memset(hist, 0, 50 * sizeof(int64_t));
    const long long total = g.nblk_y + 2 * g.nblk_c;
    for (long long b = 0; b < total; ++b) {
        int plane = 0;
        long long lb = b;
        if (lb >= g.nblk_y) { lb -= g.nblk_y; plane = 1; if (lb >= g.nblk_c) { lb -= g.nblk_c; plane = 2; } }
        const int nbx = plane ? g.nbx_c : g.nbx_y;
        const int h = plane ? g.hc : g.H, w = plane ? g.wc : g.W;
        const int stride = plane ? g.wcp : g.Wp;
        const size_t poff = plane == 0 ? 0 : (plane == 1 ? g.plane_y : g.plane_y + g.plane_c);
        const int by = (int)(lb / nbx), bx = (int)(lb % nbx);
        T v[64];
        int16_t q[64];
        BlockStats st;
        load_block<T>(fwd.data() + poff, stride, h, w, bx, by, v);
        BlockCodec<P>::run(v, q, tb, st, nullptr, nullptr);
        store_block<T>(rec.data() + poff, stride, h, w, bx, by, v);
        *bits += st.bits;
        *nnz += st.nnz;
        for (int i = 0; i < 64; ++i) {
            coeffs[b * 64 + i] = q[i];
            int hb = hist_bin(q[i]);
            if (hb >= 0) hist[hb]++;
        }
    }
    *sse_rgb = 0;
    *sse_y = 0.0;
    const T* Yr = rec.data();
    const T* Cbr = Yr + g.plane_y;
    const T* Crr = Cbr + g.plane_c;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            PixelOut o = inverse_pixel<P>(g, rgb, x, y, Y, Yr, Cbr, Crr);
            uint8_t* out = recon + ((size_t)y * W + x) * 3;
            out[0] = o.r; out[1] = o.g; out[2] = o.b;
            err_y[(size_t)y * W + x] = o.err_y;
            err_rgb[(size_t)y * W + x] = o.err_rgb;
            *sse_rgb += o.sse_rgb;
            *sse_y += o.sse_y;
        }
    return 0;
}

extern "C" int emul_roundtrip(int exact, int H, int W, int quality, int sub, int prefilter,
                              const uint8_t* rgb, uint8_t* recon, int16_t* coeffs, double* err_y,
                              double* err_rgb, uint64_t* bits, uint64_t* nnz, int64_t* hist,
                              uint64_t* sse_rgb, double* sse_y) {
    if (exact)
        return run<Exact>(H, W, quality, sub, prefilter, rgb, recon, coeffs, err_y, err_rgb, bits,
                          nnz, hist, sse_rgb, sse_y);
    return run<Fast>(H, W, quality, sub, prefilter, rgb, recon, coeffs, err_y, err_rgb, bits, nnz,
                     hist, sse_rgb, sse_y);
}

// SSIM of one window from its sums, both precisions (checked against the stand-in)
extern "C" double emul_ssim_window(int use_float, const double* x, const double* y, double shift) {
    double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
    for (int i = 0; i < 49; ++i) {
        const double a = x[i] - shift, b = y[i] - shift;
        sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
    }
    if (use_float)
        return (double)ssim_from_sums<float>((float)sx, (float)sy, (float)sxx, (float)syy,
                                             (float)sxy, (float)shift);
    return ssim_from_sums<double>(sx, sy, sxx, syy, sxy, shift);
}

// Markstein quotient used by BlockCodec<Exact> against the IEEE division, n samples per
// divisor 1..255 (returns the number of mismatches)
extern "C" long emul_division_selftest(long n_per_q) {
    unsigned long long st = 88172645463325252ULL;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
    long bad = 0;
    for (int q = 1; q <= 255; ++q) {
        const double Q = q;
        volatile double rqv = 1.0 / Q;
        const double rq = rqv;
        for (long i = 0; i < n_per_q; ++i) {
            double x;
            switch (i & 3) {
                case 0: x = (rnd() * 2 - 1) * 1100.0; break;
                case 1: x = (rnd() * 2 - 1) * 8.0; break;
                case 2: x = (floor(rnd() * 2200) - 1100 + 0.5) * q; break;          // exact ties
                default: x = nextafter((floor(rnd() * 2200) - 1100 + 0.5) * q, rnd() > 0.5 ? 1e9 : -1e9);
            }
            const double q0 = Exact::mul(x, rq);
            const double rem = Exact::fma(-q0, Q, x);
            const double got = Exact::fma(rem, rq, q0);
            if (got != Exact::div(x, Q)) ++bad;
        }
    }
    return bad;
}

// ---- chroma-aliasing demo front end (jds_alias.cuh), one call per kernel thread ----------
#include "../../jpeg_dsp_studio_b200/csrc/jds_alias.cuh"

extern "C" int emul_alias_subsample(int H, int W, int prefilter, const uint8_t* rgb, uint8_t* out) {
    const int hs = (H + 1) / 2, ws = (W + 1) / 2;
    const size_t n = (size_t)H * W;
    std::vector<float> Y(n), Cr(n), Cb(n), Crs((size_t)hs * ws), Cbs((size_t)hs * ws);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const uint8_t* p = rgb + ((size_t)y * W + x) * 3;
            alias_forward_px((float)p[0], (float)p[1], (float)p[2], x < 8 * (W / 8), Y[(size_t)y * W + x],
                             Cr[(size_t)y * W + x], Cb[(size_t)y * W + x]);
        }
    const float *cr_src = Cr.data(), *cb_src = Cb.data();
    size_t row_stride = 2 * (size_t)W;
    int col_stride = 2;
    if (prefilter) {
        for (int yo = 0; yo < hs; ++yo)
            for (int xo = 0; xo < ws; ++xo) {
                Crs[(size_t)yo * ws + xo] = alias_blur_at(Cr.data(), H, W, 2 * xo, 2 * yo);
                Cbs[(size_t)yo * ws + xo] = alias_blur_at(Cb.data(), H, W, 2 * xo, 2 * yo);
            }
        cr_src = Crs.data();
        cb_src = Cbs.data();
        row_stride = (size_t)ws;
        col_stride = 1;
    }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float cr = alias_upsample_at(cr_src, row_stride, col_stride, hs, ws, H, W, x, y);
            const float cb = alias_upsample_at(cb_src, row_stride, col_stride, hs, ws, H, W, x, y);
            alias_inverse_px(Y[(size_t)y * W + x], cr, cb, out + ((size_t)y * W + x) * 3);
        }
    return 0;
}

extern "C" void emul_alias_luma_diff(size_t n_px, const uint8_t* a, const uint8_t* b, uint8_t* luma_a,
                                     uint8_t* diff) {
    for (size_t i = 0; i < n_px; ++i) {
        luma_a[i] = alias_luma_u8(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
        for (int k = 0; k < 3; ++k) diff[3 * i + k] = alias_diff_u8(a[3 * i + k], b[3 * i + k]);
    }
}

// ---- SSIM window formula of k_ssim_strip (jds_ssim_formula.cuh) ---------------------------
#include "../../jpeg_dsp_studio_b200/csrc/jds_ssim_formula.cuh"

// n windows; sums are the centred (x - 128) fp32 window sums the kernel forms; out = SSIM
extern "C" void emul_ssim_strip_formula(int n, const float* sx, const float* sy, const float* sq,
                                        const float* sc, float* out) {
    for (int i = 0; i < n; ++i) {
        const Pair2::V r = ssim_window_half_acc<Pair2>(Pair2::splat(sx[i]), Pair2::splat(sy[i]),
                                                       Pair2::splat(sq[i]), Pair2::splat(sc[i]),
                                                       Pair2::splat(0.0f));
        out[i] = 2.0f * r.x;
    }
}

// ---------------------------------------------------------------------------------------
// entropy coder: the per-block code of jds_entropy.cu (jds_entropy_block.cuh) run block after
// block on the host - sizes by CountSink, bits by BitSink at the running bit position, the padding
// 1-bits, big-endian bytes and the 0xFF stuffing - for one scan.  Returns 0, or -1 if a value has
// no baseline code, or -2 if `cap` is too small.
// ---------------------------------------------------------------------------------------
#include "../../jpeg_dsp_studio_b200/csrc/jds_entropy_block.cuh"

extern "C" int emul_entropy_scan(const int16_t* blocks, long long n_blocks, int table_id, uint8_t* out,
                                 uint64_t cap, uint64_t* out_bytes, uint64_t* out_bits) {
    static const uint8_t zz[64] = JDS_ZIGZAG_TABLE;
    const HuffPacked hp = make_packed();
    const uint32_t* dc = hp.dc[table_id ? 1 : 0];
    const uint32_t* ac = hp.ac[table_id ? 1 : 0];
    // pass 1: sizes (what k_entropy_block_bits + k_entropy_layout produce)
    std::vector<uint64_t> start((size_t)n_blocks + 1, 0);
    int pred = 0;
    for (long long b = 0; b < n_blocks; ++b) {
        CountSink cs;
        if (!walk_block(blocks + 64 * b, pred, dc, ac, zz, cs)) return -1;
        pred = blocks[64 * b];
        start[(size_t)b + 1] = start[(size_t)b] + cs.bits;
    }
    const uint64_t bits = start[(size_t)n_blocks];
    const int pad = (int)((8 - (bits & 7)) & 7);
    std::vector<uint32_t> words((size_t)((bits + pad + 31) / 32) + 1, 0u);
    // pass 2: every block at its own offset (k_entropy_pack), in REVERSE order to show that the
    // result does not depend on the order in which blocks write
    for (long long b = n_blocks - 1; b >= 0; --b) {
        BitSink sink(words.data(), (unsigned int)start[(size_t)b]);
        walk_block(blocks + 64 * b, b ? (int)blocks[64 * (b - 1)] : 0, dc, ac, zz, sink);
        if (b == n_blocks - 1 && pad) sink.put((1u << pad) - 1u, pad);
        sink.finish();
    }
    // bytes, first bit = MSB of byte 0, 0x00 after every 0xFF (k_stuff_*)
    const uint64_t nbytes = (bits + pad) / 8;
    uint64_t at = 0;
    for (uint64_t i = 0; i < nbytes; ++i) {
        const uint8_t v = (uint8_t)(words[(size_t)(i >> 2)] >> (24 - 8 * (i & 3)));
        if (at + 2 > cap) return -2;
        out[at++] = v;
        if (v == 0xFF) out[at++] = 0;
    }
    *out_bytes = at;
    *out_bits = bits;
    return 0;
}
