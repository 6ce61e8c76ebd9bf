// jds_stages.cuh - the per-work-item bodies of the staged pipeline.
//
// Each function is the work of ONE thread of the corresponding kernel in
// jds_kernels.cu (a chroma cell, an 8x8 block, a pixel).  They are __host__
// __device__ so tests/emul can loop over them on the CPU and compare with the
// oracle; the product only calls them from kernels.
#pragma once
#include "jds_internal.cuh"

namespace jds {

template <class P>
JDS_HD void load_rgb(const uint8_t* __restrict__ rgb, int W, int y, int x,
                     typename P::T& r, typename P::T& g, typename P::T& b) {
    typedef typename P::T T;
    const uint8_t* p = rgb + ((size_t)y * W + x) * 3;
    r = T(p[0]);
    g = T(p[1]);
    b = T(p[2]);
}

// ------------------------------------------------------------------------------
// Stage 1: RGB -> Y plane + decimated Cb/Cr planes.  One call per chroma cell
// (1x1, 2x1 or 2x2 luma pixels).
//   engines/color_space.py:8-14 (rgb_to_ycbcr), :27-53 (subsample_chroma)
// ------------------------------------------------------------------------------
template <class P, int SUB, bool PF>
JDS_HD void forward_cell(const Geom& g, const uint8_t* __restrict__ rgb, int cx, int cy,
                         typename P::T* __restrict__ Yp, typename P::T* __restrict__ Cbp,
                         typename P::T* __restrict__ Crp) {
    typedef typename P::T T;
    constexpr int CW = (SUB == 0) ? 1 : 2;          // cell width in luma pixels
    constexpr int CH = (SUB == 2) ? 2 : 1;          // cell height
    const int x0 = cx * CW, y0 = cy * CH;
    T cb[CH][CW], cr[CH][CW];
#pragma unroll
    for (int dy = 0; dy < CH; ++dy)
#pragma unroll
        for (int dx = 0; dx < CW; ++dx) {
            T r, gg, b, yy;
            load_rgb<P>(rgb, g.W, y0 + dy, x0 + dx, r, gg, b);
            if (Yp) {
                rgb_to_ycbcr<P>(r, gg, b, yy, cb[dy][dx], cr[dy][dx]);
                Yp[(size_t)(y0 + dy) * g.Wp + (x0 + dx)] = yy;
            } else {                                  // chroma only (fused exact path: Y never
                rgb_to_cbcr<P>(r, gg, b, cb[dy][dx], cr[dy][dx]);      // leaves the luma kernel)
            }
        }
    if (SUB != 0 && PF) {
        // cv2.GaussianBlur(3x3, sigma 0.75) of the full-resolution chroma at the
        // cell's pixels (A2): row pass on rows y0-1 .. y0+CH, then column pass.
        constexpr int NR = CH + 2;
        T rb[NR][CW], rr[NR][CW];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int yy = reflect101(y0 - 1 + i, g.H);
            T vb[CW + 2], vr[CW + 2];
#pragma unroll
            for (int j = 0; j < CW + 2; ++j) {
                const int xx = reflect101(x0 - 1 + j, g.W);
                T r, gg, b;
                load_rgb<P>(rgb, g.W, yy, xx, r, gg, b);
                rgb_to_cbcr<P>(r, gg, b, vb[j], vr[j]);
            }
#pragma unroll
            for (int dx = 0; dx < CW; ++dx) {
                const bool tail = (x0 + dx) >= g.W4;
                rb[i][dx] = blur_row<P>(vb[dx], vb[dx + 1], vb[dx + 2], tail);
                rr[i][dx] = blur_row<P>(vr[dx], vr[dx + 1], vr[dx + 2], tail);
            }
        }
#pragma unroll
        for (int dy = 0; dy < CH; ++dy)
#pragma unroll
            for (int dx = 0; dx < CW; ++dx) {
                cb[dy][dx] = blur_col<P>(rb[dy][dx], rb[dy + 1][dx], rb[dy + 2][dx]);
                cr[dy][dx] = blur_col<P>(rr[dy][dx], rr[dy + 1][dx], rr[dy + 2][dx]);
            }
    }
    T ob, orr;
    if (SUB == 0) {
        ob = cb[0][0];
        orr = cr[0][0];
    } else if (SUB == 1) {                            // A3, 4:2:2: (a+b)*0.5
        ob = P::mul(P::add(cb[0][0], cb[0][CW - 1]), T(0.5));
        orr = P::mul(P::add(cr[0][0], cr[0][CW - 1]), T(0.5));
    } else {                                          // A3, 4:2:0: (((a+b)+c)+d)*0.25
        ob = P::mul(P::add(P::add(P::add(cb[0][0], cb[0][CW - 1]), cb[CH - 1][0]),
                           cb[CH - 1][CW - 1]), T(0.25));
        orr = P::mul(P::add(P::add(P::add(cr[0][0], cr[0][CW - 1]), cr[CH - 1][0]),
                            cr[CH - 1][CW - 1]), T(0.25));
    }
    Cbp[(size_t)cy * g.wcp + cx] = ob;
    Crp[(size_t)cy * g.wcp + cx] = orr;
}

// ------------------------------------------------------------------------------
// Stage 1, general geometry: an odd width (or an odd height under 4:2:0) makes
// cv2.resize(INTER_AREA) (engines/color_space.py:44-49) leave its integer-factor path and
// weight every source sample by its overlap with the destination cell
// (OpenCV imgproc/resize.cpp computeResizeAreaTab + ResizeArea_Invoker<double,double>).
// ------------------------------------------------------------------------------
struct AreaSpan {
    int s0;            // first source index; the taps are s0 .. s0+n-1
    int n;
    float al, af, ar;  // weights of a partial first tap, of the whole taps, of a partial last
                       // tap - rounded to float like OpenCV's DecimateAlpha::alpha
    bool left, right;  // is the first / last tap a partial one?
    JDS_HD float weight(int i) const {
        return (i == 0 && left) ? al : ((i == n - 1 && right) ? ar : af);
    }
};

// Taps of destination index d when `ssize` samples shrink to `dsize`.  Every operation is an
// individually rounded fp64 one, in OpenCV's order (both policies: taps are geometry).
JDS_HD void area_span(int ssize, int dsize, int d, AreaSpan& t) {
    typedef Exact E;
    const double scale = E::div((double)ssize, (double)dsize);
    const double fsx1 = E::mul((double)d, scale);
    const double fsx2 = E::add(fsx1, scale);
    const double rest = E::sub((double)ssize, fsx1);
    const double cell = scale < rest ? scale : rest;
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    if (sx2 > ssize - 1) sx2 = ssize - 1;
    if (sx1 > sx2) sx1 = sx2;
    const double lw = E::sub((double)sx1, fsx1);
    const double rw = E::sub(fsx2, (double)sx2);
    t.left = lw > 1e-3;
    t.right = rw > 1e-3;
    t.s0 = t.left ? sx1 - 1 : sx1;
    t.n = (sx2 - sx1) + (t.left ? 1 : 0) + (t.right ? 1 : 0);
    t.al = (float)E::div(lw, cell);
    t.af = (float)E::div(1.0, cell);
    double m = rw < 1.0 ? rw : 1.0;
    if (cell < m) m = cell;
    t.ar = (float)E::div(m, cell);
}

// full-resolution Cb, Cr of pixel (y, x), after the optional 3x3 prefilter (A2)
template <class P, bool PF>
JDS_HD void chroma_at(const Geom& g, const uint8_t* __restrict__ rgb, int y, int x,
                      typename P::T& cb, typename P::T& cr) {
    typedef typename P::T T;
    if (!PF) {
        T r, gg, b;
        load_rgb<P>(rgb, g.W, y, x, r, gg, b);
        rgb_to_cbcr<P>(r, gg, b, cb, cr);
        return;
    }
    T rb[3], rr[3];
    const bool tail = x >= g.W4;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int yy = reflect101(y - 1 + i, g.H);
        T vb[3], vr[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            T r, gg, b;
            load_rgb<P>(rgb, g.W, yy, reflect101(x - 1 + j, g.W), r, gg, b);
            rgb_to_cbcr<P>(r, gg, b, vb[j], vr[j]);
        }
        rb[i] = blur_row<P>(vb[0], vb[1], vb[2], tail);
        rr[i] = blur_row<P>(vr[0], vr[1], vr[2], tail);
    }
    cb = blur_col<P>(rb[0], rb[1], rb[2]);
    cr = blur_col<P>(rr[0], rr[1], rr[2]);
}

// one luma sample (general geometry: the chroma cells do not tile the luma plane)
template <class P>
JDS_HD void forward_luma(const Geom& g, const uint8_t* __restrict__ rgb, int x, int y,
                         typename P::T* __restrict__ Yp) {
    typename P::T r, gg, b;
    load_rgb<P>(rgb, g.W, y, x, r, gg, b);
    Yp[(size_t)y * g.Wp + x] = luma601<P>(r, gg, b);
}

// one decimated chroma sample by area weighting: per source row buf = buf + S*alpha over
// the x taps in order, then sum = beta*buf (first row) / sum + beta*buf (no FMA)
template <class P, bool PF>
JDS_HD void forward_chroma_area(const Geom& g, const uint8_t* __restrict__ rgb, int cx, int cy,
                                typename P::T* __restrict__ Cbp, typename P::T* __restrict__ Crp) {
    typedef typename P::T T;
    AreaSpan tx, ty;
    area_span(g.W, g.wc, cx, tx);
    area_span(g.H, g.hc, cy, ty);
    T sum_b = T(0), sum_r = T(0);
    for (int j = 0; j < ty.n; ++j) {
        T buf_b = T(0), buf_r = T(0);
        for (int i = 0; i < tx.n; ++i) {
            T cb, cr;
            chroma_at<P, PF>(g, rgb, ty.s0 + j, tx.s0 + i, cb, cr);
            buf_b = P::add(buf_b, P::mul(cb, T(tx.weight(i))));
            buf_r = P::add(buf_r, P::mul(cr, T(tx.weight(i))));
        }
        const T beta = T(ty.weight(j));
        sum_b = j == 0 ? P::mul(beta, buf_b) : P::add(sum_b, P::mul(beta, buf_b));
        sum_r = j == 0 ? P::mul(beta, buf_r) : P::add(sum_r, P::mul(beta, buf_r));
    }
    Cbp[(size_t)cy * g.wcp + cx] = sum_b;
    Crp[(size_t)cy * g.wcp + cx] = sum_r;
}

// ------------------------------------------------------------------------------
// Stage 2: one 8x8 block: (reflect-padded) load, -128, DCT, quantise, bit model,
// dequantise, IDCT, +128, clip.   engines/pipeline.py:47-82 inner loops,
// engines/dct_engine.py:17-27, engines/quantizer.py:22-29.
// ------------------------------------------------------------------------------
struct BlockStats {
    unsigned int bits;
    unsigned int nnz;
};

template <class P>
struct BlockCodec;

template <>
struct BlockCodec<Exact> {
    // v: 64 samples (row-major) in, reconstructed samples out; q: quantised out.
    //
    // The arithmetic is SURVEY Appendix A5-A7 bit for bit, with the transforms' power-of-two
    // multiplications deferred (dct8_ref_unscaled / idct8_ref_unscaled, jds_math.cuh): the
    // forward coefficients stay multiplied by 2^shift(i) (exact_coeff_shift), which the
    // quantiser absorbs - the correctly rounded quotient of a scaled value is the scaled
    // quotient, the rounding constant is scaled with it, and the dequantiser table dqx carries
    // 2^-shift together with the 1/16 of the inverse's first axis.
    template <int I>
    static JDS_HD void quant_one(double* v, int16_t* q, const QTables& tb, unsigned int* esum,
                                 unsigned int* nnz, double* dct_out, double* deq_out) {
        typedef Exact P;
        constexpr int SH = exact_coeff_shift(I);
        if (dct_out) dct_out[I] = P::mul(v[I], 1.0 / (double)(1 << SH));
        // X / Q correctly rounded without the division subroutine (Markstein): with
        // y = RN(1/Q), q0 = RN(X y), r = X - q0 Q (exact in one FMA), RN(q0 + r y) is the
        // IEEE quotient for every integer Q in 1..255 (102 M cases incl. near-ties checked
        // against x/Q in tests/emul; tests/test_device_math_on_cpu.py)
        const double q0 = P::mul(v[I], tb.rq[I]);
        const double rem = P::fma(-q0, tb.q[I], v[I]);
        const RoundedQuotient rq = round_half_even<SH>(P::fma(rem, tb.rq[I], q0));
        q[I] = (int16_t)rq.ivalue;
        esum[SH] += (unsigned int)rq.expo;
        nnz[SH] += ((unsigned int)rq.expo + 2047u) >> 11;       // 1 for every non-zero value
        v[I] = P::mul(rq.value, tb.dqx[I]);     // = int16 * Q / 16: never -0.0 (quantizer.py:29)
        if (deq_out) deq_out[I] = P::mul(v[I], 16.0);
    }
    template <int I0>
    static JDS_HD void quant_row(double* v, int16_t* q, const QTables& tb, unsigned int* esum,
                                 unsigned int* nnz, double* dct_out, double* deq_out) {
        quant_one<I0 + 0>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_one<I0 + 1>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_one<I0 + 2>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_one<I0 + 3>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_one<I0 + 4>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_one<I0 + 5>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_one<I0 + 6>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_one<I0 + 7>(v, q, tb, esum, nnz, dct_out, deq_out);
    }

    static JDS_HD void run(double* v, int16_t* q, const QTables& tb, BlockStats& st,
                           double* dct_out, double* deq_out) {
        typedef Exact P;
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = P::sub(v[i], 128.0);
        // axis 0 (down each column) first, then axis 1 (A5 '2-D')
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double t[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) t[r] = v[r * 8 + c];
            dct8_ref_unscaled<P>(t);
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r * 8 + c] = t[r];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) dct8_ref_unscaled<P>(v + r * 8);
        // bit model (utils/metrics.py:75-79): 6 + ceil(log2(|v|+1)) + 1 = 7 + bit_length(|v|)
        // per non-zero value; bit_length comes from the exponent field of the rounded (still
        // scaled) quotient: biased exponent - 1022 - shift, summed per shift class
        unsigned int esum[5] = {0, 0, 0, 0, 0}, nnz[5] = {0, 0, 0, 0, 0};
        quant_row<0>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_row<8>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_row<16>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_row<24>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_row<32>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_row<40>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_row<48>(v, q, tb, esum, nnz, dct_out, deq_out);
        quant_row<56>(v, q, tb, esum, nnz, dct_out, deq_out);
        unsigned int bits = 0, n = 0;
#pragma unroll
        for (int sh = 2; sh <= 4; ++sh) {       // the shifts that occur: 2, 3, 4
            bits += esum[sh] - (1015u + (unsigned)sh) * nnz[sh];
            n += nnz[sh];
        }
        st.bits = bits;
        st.nnz = n;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double t[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) t[r] = v[r * 8 + c];
            idct8_ref_unscaled<P>(t);
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r * 8 + c] = t[r];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) idct8_ref_unscaled<P>(v + r * 8);
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = P::clamp255(P::add(v[i], 128.0));
    }
};

template <>
struct BlockCodec<Fast> {
    static JDS_HD void run(float* v, int16_t* q, const QTables& tb, BlockStats& st,
                           float* dct_out, float* deq_out) {
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] -= 128.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float t[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) t[r] = v[r * 8 + c];
            dct8_aan(t);
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r * 8 + c] = t[r];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) dct8_aan(v + r * 8);
        unsigned int bits = 0, nnz = 0;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            const float qv = rintf(v[i] * tb.fq[i]);
            if (dct_out) dct_out[i] = qv;      // fast mode: diagnostic only
            const int qi = (int)qv;
            q[i] = (int16_t)qi;
            bits += coeff_bits(qi);
            nnz += (qi != 0);
            v[i] = qv * tb.dq[i];
            if (deq_out) deq_out[i] = v[i];
        }
        st.bits = bits;
        st.nnz = nnz;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float t[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) t[r] = v[r * 8 + c];
            idct8_aan(t);
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r * 8 + c] = t[r];
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) idct8_aan(v + r * 8);
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = fminf(fmaxf(v[i] + 128.0f, 0.0f), 255.0f);
    }
};

// load the 64 samples of block (bx, by) of a plane whose valid size is h x w and
// whose row stride is `stride`; positions beyond the valid size are np.pad reflect.
template <class T>
JDS_HD void load_block(const T* __restrict__ plane, int stride, int h, int w, int bx, int by,
                       T* v) {
    const int x0 = bx * 8, y0 = by * 8;
    if (x0 + 8 <= w && y0 + 8 <= h) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) v[r * 8 + c] = plane[(size_t)(y0 + r) * stride + x0 + c];
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int yy = reflect_index(y0 + r, h);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                v[r * 8 + c] = plane[(size_t)yy * stride + reflect_index(x0 + c, w)];
        }
    }
}

template <class T>
JDS_HD void store_block(T* __restrict__ plane, int stride, int h, int w, int bx, int by,
                        const T* v) {
    const int x0 = bx * 8, y0 = by * 8;
    if (x0 + 8 <= w && y0 + 8 <= h) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) plane[(size_t)(y0 + r) * stride + x0 + c] = v[r * 8 + c];
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (y0 + r < h && x0 + c < w)
                    plane[(size_t)(y0 + r) * stride + x0 + c] = v[r * 8 + c];
    }
}

// ------------------------------------------------------------------------------
// Stage 3: one output pixel: chroma upsample (A8), YCbCr -> RGB (A9), clamp,
// truncate.   engines/color_space.py:56-66, :17-24; engines/pipeline.py:88-95
// ------------------------------------------------------------------------------
template <class P>
JDS_HD void upsample_taps(int i, int n_src, double step, int& i0, int& i1, typename P::T& f) {
    // IPP forms the source coordinate with one FMA (visible only when the factor is not
    // exactly 2, i.e. odd sizes)
    double ff = Exact::fma((double)i + 0.5, step, -0.5);
    double s = floor(ff);
    ff -= s;
    int si = (int)s;
    i0 = si < 0 ? 0 : (si > n_src - 1 ? n_src - 1 : si);
    i1 = si + 1 < 0 ? 0 : (si + 1 > n_src - 1 ? n_src - 1 : si + 1);
    f = (typename P::T)ff;
}

template <class P>
JDS_HD typename P::T upsample_sample(const typename P::T* __restrict__ S, int stride, int sub,
                                     int x0, int x1, typename P::T fx,
                                     int y0, int y1, typename P::T fy) {
    typedef typename P::T T;
    const T a0 = S[(size_t)y0 * stride + x0], a1 = S[(size_t)y0 * stride + x1];
    const T t0 = P::fma(P::sub(a1, a0), fx, a0);
    if (sub == 1) return t0;                          // 4:2:2: rows are not resampled
    const T b0 = S[(size_t)y1 * stride + x0], b1 = S[(size_t)y1 * stride + x1];
    const T t1 = P::fma(P::sub(b1, b0), fx, b0);
    return P::fma(P::sub(t1, t0), fy, t0);
}

struct PixelOut {
    uint8_t r, g, b;
    double err_y, err_rgb;      // IntermediateData.error_map_* (pipeline.py:119-121)
    unsigned int sse_rgb;       // (dr^2 + dg^2 + db^2) on the uint8 values
    double sse_y;               // (Yo - Yr)^2 on BT.601 Y of the uint8 values (metrics.py:17-20)
};

template <class P>
JDS_HD PixelOut inverse_pixel(const Geom& g, const uint8_t* __restrict__ rgb, int x, int y,
                              const typename P::T* __restrict__ Yf,
                              const typename P::T* __restrict__ Yr,
                              const typename P::T* __restrict__ Cbr,
                              const typename P::T* __restrict__ Crr) {
    typedef typename P::T T;
    PixelOut o;
    const T yv = Yr[(size_t)y * g.Wp + x];
    T cb, cr;
    if (g.sub == 0) {
        cb = Cbr[(size_t)y * g.wcp + x];
        cr = Crr[(size_t)y * g.wcp + x];
    } else {
        int x0, x1, y0, y1;
        T fx, fy;
        upsample_taps<P>(x, g.wc, g.sx, x0, x1, fx);
        upsample_taps<P>(y, g.hc, g.sy, y0, y1, fy);
        cb = upsample_sample<P>(Cbr, g.wcp, g.sub, x0, x1, fx, y0, y1, fy);
        cr = upsample_sample<P>(Crr, g.wcp, g.sub, x0, x1, fx, y0, y1, fy);
    }
    T rf, gf, bf;
    ycbcr_to_rgb<P>(yv, cb, cr, rf, gf, bf);
    o.r = (uint8_t)(int)rf;                           // .astype(uint8): truncation
    o.g = (uint8_t)(int)gf;
    o.b = (uint8_t)(int)bf;
    const uint8_t* p = rgb + ((size_t)y * g.W + x) * 3;
    const int dr = (int)p[0] - (int)o.r, dg = (int)p[1] - (int)o.g, db = (int)p[2] - (int)o.b;
    o.sse_rgb = (unsigned)(dr * dr + dg * dg + db * db);
    // Y of the uint8 images, fp64 un-fused like NumPy (metrics.py:17-18)
    const double yo = luma601<Exact>((double)p[0], (double)p[1], (double)p[2]);
    const double yr8 = luma601<Exact>((double)o.r, (double)o.g, (double)o.b);
    const double dyv = yo - yr8;
    o.sse_y = dyv * dyv;
    o.err_y = (double)P::abs_(P::sub(Yf[(size_t)y * g.Wp + x], yv));
    const T er = P::abs_(P::sub(T(p[0]), rf));
    const T eg = P::abs_(P::sub(T(p[1]), gf));
    const T eb = P::abs_(P::sub(T(p[2]), bf));
    o.err_rgb = (double)P::div(P::add(P::add(er, eg), eb), T(3.0));
    return o;
}

// ------------------------------------------------------------------------------
// Stage 4: SSIM of one 7x7 window from its five sums (skimage
// _structural_similarity.py: uniform_filter means, sample covariance, K1=.01, K2=.03,
// data_range 255).  Sums are over the 49 samples.
// ------------------------------------------------------------------------------
// The samples may have been shifted by -shift before summing (variances and
// covariance are shift invariant; only the means need the shift back).
template <class T>
JDS_HD T ssim_from_sums(T sx, T sy, T sxx, T syy, T sxy, T shift) {
    const T inv = T(1.0 / 49.0);
    const T cov_norm = T(49.0 / 48.0);
    const T C1 = T(0.01 * 255.0) * T(0.01 * 255.0);
    const T C2 = T(0.03 * 255.0) * T(0.03 * 255.0);
    const T uxs = sx * inv, uys = sy * inv;
    const T uxx = sxx * inv, uyy = syy * inv, uxy = sxy * inv;
    const T vx = cov_norm * (uxx - uxs * uxs);
    const T vy = cov_norm * (uyy - uys * uys);
    const T vxy = cov_norm * (uxy - uxs * uys);
    const T ux = uxs + shift, uy = uys + shift;
    const T A1 = T(2) * ux * uy + C1, A2 = T(2) * vxy + C2;
    const T B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
    return (A1 * A2) / (B1 * B2);
}

}  // namespace jds
