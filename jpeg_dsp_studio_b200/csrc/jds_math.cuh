// jds_math.cuh - arithmetic of the round trip, shared by every kernel.
//
// Two arithmetic policies:
//   Exact  fp64, every operation individually rounded and never contracted, in the
//          operation order of the libraries the reference calls (SURVEY.md
//          Appendix A): NumPy colour math (engines/color_space.py:8-24), ducc0's
//          8-point DCT as executed by scipy.fft.dctn/idctn (engines/dct_engine.py:7-14),
//          OpenCV's fp64 GaussianBlur / INTER_AREA / IPP INTER_LINEAR
//          (engines/color_space.py:39-49,64-65).
//   Fast   fp32, contraction allowed, scaled AAN butterflies with the scale factors
//          folded into the quantiser tables.
//
// Everything here is __host__ __device__ so that tests/emul can run the very same
// source on the CPU against the oracle before any GPU time is spent; the product
// library only ever calls it from kernels.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define JDS_HD __host__ __device__ __forceinline__
#else
#define JDS_HD inline
#endif

namespace jds {

// ------------------------------------------------------------------------------
// policies
// ------------------------------------------------------------------------------
struct Exact {
    typedef double T;
    static constexpr bool kExact = true;
#if defined(__CUDA_ARCH__)
    static JDS_HD T add(T a, T b) { return __dadd_rn(a, b); }
    static JDS_HD T sub(T a, T b) { return __dsub_rn(a, b); }
    static JDS_HD T mul(T a, T b) { return __dmul_rn(a, b); }
    static JDS_HD T div(T a, T b) { return __ddiv_rn(a, b); }
    static JDS_HD T fma(T a, T b, T c) { return __fma_rn(a, b, c); }
#else
    // host build (tests/emul): compiled with -ffp-contract=off and no FMA ISA
    static JDS_HD T add(T a, T b) { volatile T r = a + b; return r; }
    static JDS_HD T sub(T a, T b) { volatile T r = a - b; return r; }
    static JDS_HD T mul(T a, T b) { volatile T r = a * b; return r; }
    static JDS_HD T div(T a, T b) { volatile T r = a / b; return r; }
    static JDS_HD T fma(T a, T b, T c) { return ::fma(a, b, c); }
#endif
    static JDS_HD T rint_(T a) { return ::rint(a); }
    static JDS_HD T abs_(T a) { return ::fabs(a); }
    // np.clip(a, 0, 255) for finite a as two compares and selects (fmin / fmax in fp64 expand
    // to ~14 instructions each for their NaN rules: a quarter of the exact luma kernel's
    // instructions before this); -0.0 and negative values give +0.0
#if defined(__CUDA_ARCH__)
    // on the device the two comparisons read the high word as an integer (sign bit: negative
    // or -0.0; >= 0x406FE000: 255.0 and above), which keeps them off the fp64 pipe - the pipe
    // that bounds the exact kernels; same results as the floating-point form below
    static JDS_HD T clamp255(T a) {
        const int hi = __double2hiint(a);
        return hi < 0 ? 0.0 : (hi >= 0x406FE000 ? 255.0 : a);
    }
#else
    static JDS_HD T clamp255(T a) { return a > 0.0 ? (a > 255.0 ? 255.0 : a) : 0.0; }
#endif
};

// np.round(x).astype(int16) of a quotient |x| < 2^31 (engines/quantizer.py:24): adding
// 1.5 * 2^52 rounds to the nearest integer, ties to even (one IEEE addition), and leaves that
// integer in the low word of the sum; subtracting the constant again gives it as a double -
// +0.0, never -0.0, like the reference's int16 -> float64 cast (quantizer.py:29).
struct RoundedQuotient {
    double value;   // the rounded quotient as fp64 (times 2^SHIFT)
    int ivalue;     // the rounded quotient as an integer
    int expo;       // biased exponent field of `value` (0 for zero): bit_length(|v|) = expo - 1022 - SHIFT
};
// `x` holds the quotient TIMES 2^SHIFT (the block codec keeps a power-of-two scale in its
// coefficients).  Adding 1.5 * 2^(52+SHIFT) rounds x to a multiple of 2^SHIFT, i.e. the true
// quotient to an integer (ties to even; the scaling is exact), and that integer again sits in
// the low word of the sum.  SHIFT = 0 is plain np.round.
template <int SHIFT>
JDS_HD RoundedQuotient round_half_even(double x) {
    RoundedQuotient r;
    constexpr double MAGIC = 6755399441055744.0 * (double)(1ull << SHIFT);
    const double t = Exact::add(x, MAGIC);
    r.value = Exact::sub(t, MAGIC);
#if defined(__CUDA_ARCH__)
    r.ivalue = __double2loint(t);
    r.expo = (__double2hiint(r.value) >> 20) & 0x7FF;
#else
    unsigned long long tb, vb;
    memcpy(&tb, &t, 8);
    memcpy(&vb, &r.value, 8);
    r.ivalue = (int)(unsigned int)(tb & 0xFFFFFFFFull);
    r.expo = (int)((vb >> 52) & 0x7FF);
#endif
    return r;
}

struct Fast {
    typedef float T;
    static constexpr bool kExact = false;
    static JDS_HD T add(T a, T b) { return a + b; }
    static JDS_HD T sub(T a, T b) { return a - b; }
    static JDS_HD T mul(T a, T b) { return a * b; }
    static JDS_HD T div(T a, T b) { return a / b; }
    static JDS_HD T fma(T a, T b, T c) { return ::fmaf(a, b, c); }
    static JDS_HD T rint_(T a) { return ::rintf(a); }
    static JDS_HD T abs_(T a) { return ::fabsf(a); }
    static JDS_HD T clamp255(T a) { return ::fminf(::fmaxf(a, 0.0f), 255.0f); }
};

// ------------------------------------------------------------------------------
// constants
// ------------------------------------------------------------------------------
// ducc0 UnityRoots twiddles for N=8: libm-derived, NOT correctly rounded cosines.
#define JDS_T0 0x1.f6297cff75cb0p-1
#define JDS_T1 0x1.d906bcf328d46p-1
#define JDS_T2 0x1.a9b66290ea1a3p-1
#define JDS_T3 0x1.6a09e667f3bccp-1
#define JDS_T4 0x1.1c73b39ae68c8p-1
#define JDS_T5 0x1.87de2a6aea963p-2
#define JDS_T6 0x1.8f8b83c69a60ap-3
#define JDS_WR 0x1.6a09e667f3bccp-1
#define JDS_WI 0x1.6a09e667f3bcdp-1
#define JDS_S2 0x1.6a09e667f3bcdp+0
#define JDS_HALF_S2 0x1.6a09e667f3bcdp-1
// cv2.getGaussianKernel(3, 0.75), fp64
#define JDS_KE 0x1.ce0cac8ce5377p-3
#define JDS_KC 0x1.18f9a9b98d643p-1

// ------------------------------------------------------------------------------
// index helpers
// ------------------------------------------------------------------------------
// np.pad(mode='reflect') source index (engines/block_processor.py:13): no edge repeat,
// periodic for pads longer than the axis.
JDS_HD int reflect_index(int i, int n) {
    if (n == 1) return 0;
    int period = 2 * (n - 1);
    i = i % period;
    return i < n ? i : period - i;
}
// cv2 BORDER_REFLECT_101 for a 3-tap filter (offsets -1 / n only)
JDS_HD int reflect101(int i, int n) {
    if (n == 1) return 0;               // OpenCV's borderInterpolate: a 1-sample axis repeats it
    if (i < 0) return -i;
    if (i >= n) return 2 * (n - 1) - i;
    return i;
}

// ------------------------------------------------------------------------------
// colour (A1, A9)
// ------------------------------------------------------------------------------
template <class P>
JDS_HD typename P::T luma601(typename P::T r, typename P::T g, typename P::T b) {
    typedef typename P::T T;
    return P::add(P::add(P::mul(T(0.299), r), P::mul(T(0.587), g)), P::mul(T(0.114), b));
}
template <class P>
JDS_HD void rgb_to_ycbcr(typename P::T r, typename P::T g, typename P::T b,
                         typename P::T& y, typename P::T& cb, typename P::T& cr) {
    typedef typename P::T T;
    y = luma601<P>(r, g, b);
    cb = P::add(P::add(P::sub(P::mul(T(-0.168736), r), P::mul(T(0.331264), g)),
                       P::mul(T(0.5), b)), T(128.0));
    cr = P::add(P::sub(P::sub(P::mul(T(0.5), r), P::mul(T(0.418688), g)),
                       P::mul(T(0.081312), b)), T(128.0));
}
template <class P>
JDS_HD void rgb_to_cbcr(typename P::T r, typename P::T g, typename P::T b,
                        typename P::T& cb, typename P::T& cr) {
    typedef typename P::T T;
    cb = P::add(P::add(P::sub(P::mul(T(-0.168736), r), P::mul(T(0.331264), g)),
                       P::mul(T(0.5), b)), T(128.0));
    cr = P::add(P::sub(P::sub(P::mul(T(0.5), r), P::mul(T(0.418688), g)),
                       P::mul(T(0.081312), b)), T(128.0));
}
// engines/color_space.py:20-23 followed by pipeline.py:95 - result clamped, NOT rounded
template <class P>
JDS_HD void ycbcr_to_rgb(typename P::T y, typename P::T cb, typename P::T cr,
                         typename P::T& r, typename P::T& g, typename P::T& b) {
    typedef typename P::T T;
    T cbs = P::sub(cb, T(128.0));
    T crs = P::sub(cr, T(128.0));
    r = P::clamp255(P::add(y, P::mul(T(1.402), crs)));
    g = P::clamp255(P::sub(P::sub(y, P::mul(T(0.344136), cbs)), P::mul(T(0.714136), crs)));
    b = P::clamp255(P::add(y, P::mul(T(1.772), cbs)));
}

// ------------------------------------------------------------------------------
// prefilter taps (A2)
// ------------------------------------------------------------------------------
// horizontal pass of cv2.GaussianBlur(3x3, 0.75) on fp64: FMA form in the vectorised
// body (columns < 4*floor(W/4)), plain form in the scalar tail.
template <class P>
JDS_HD typename P::T blur_row(typename P::T xm, typename P::T x0, typename P::T xp, bool tail) {
    typedef typename P::T T;
    if (P::kExact && tail)
        return P::add(P::add(P::mul(T(JDS_KE), xm), P::mul(T(JDS_KC), x0)), P::mul(T(JDS_KE), xp));
    return P::fma(T(JDS_KE), xp, P::fma(T(JDS_KC), x0, P::mul(T(JDS_KE), xm)));
}
template <class P>
JDS_HD typename P::T blur_col(typename P::T rm, typename P::T r0, typename P::T rp) {
    typedef typename P::T T;
    return P::add(P::mul(T(JDS_KC), r0), P::mul(T(JDS_KE), P::add(rp, rm)));
}

// ------------------------------------------------------------------------------
// 8-point transforms, exact (A5, A6): in place on c[0..7]; f is the power-of-two
// scale ducc0 applies inside the pass (1/16 on the first axis, 1 on the second).
// ------------------------------------------------------------------------------
template <class P>
JDS_HD void dct8_ref(typename P::T* c, typename P::T f) {
    typedef typename P::T T;
    const T TW[7] = {T(JDS_T0), T(JDS_T1), T(JDS_T2), T(JDS_T3), T(JDS_T4), T(JDS_T5), T(JDS_T6)};
    c[0] = P::mul(c[0], T(2.0));
    c[7] = P::mul(c[7], T(2.0));
#pragma unroll
    for (int k = 1; k <= 5; k += 2) {
        T a = c[k + 1], b = c[k];
        c[k + 1] = P::sub(a, b);
        c[k] = P::add(a, b);
    }
    T h[8];
    h[0] = P::add(c[0], c[7]);
    h[4] = P::sub(c[0], c[7]);
    h[3] = P::mul(T(2.0), c[3]);
    h[7] = P::mul(T(-2.0), c[4]);
    h[1] = P::add(c[1], c[5]);
    T tr = P::sub(c[1], c[5]);
    T ti = P::add(c[2], c[6]);
    h[2] = P::sub(c[2], c[6]);
    h[6] = P::add(P::mul(T(JDS_WR), ti), P::mul(T(JDS_WI), tr));
    h[5] = P::sub(P::mul(T(JDS_WR), tr), P::mul(T(JDS_WI), ti));
    T r[8];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int b = 4 * k;
        T u = P::add(h[b], h[b + 3]);
        T v = P::sub(h[b], h[b + 3]);
        T p = P::mul(T(2.0), h[b + 1]);
        T q = P::mul(T(2.0), h[b + 2]);
        r[k] = P::add(u, p);
        r[k + 4] = P::sub(u, p);
        r[k + 6] = P::add(v, q);
        r[k + 2] = P::sub(v, q);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = P::mul(r[i], f);
    c[0] = P::mul(r[0], T(JDS_HALF_S2));
#pragma unroll
    for (int k = 1; k <= 3; ++k) {
        const int kc = 8 - k;
        T t1 = P::add(P::mul(TW[k - 1], r[kc]), P::mul(TW[kc - 1], r[k]));
        T t2 = P::sub(P::mul(TW[k - 1], r[k]), P::mul(TW[kc - 1], r[kc]));
        c[k] = P::mul(T(0.5), P::add(t1, t2));
        c[kc] = P::mul(T(0.5), P::sub(t1, t2));
    }
    c[4] = P::mul(r[4], TW[3]);
}

template <class P>
JDS_HD void idct8_ref(typename P::T* c, typename P::T f) {
    typedef typename P::T T;
    const T TW[7] = {T(JDS_T0), T(JDS_T1), T(JDS_T2), T(JDS_T3), T(JDS_T4), T(JDS_T5), T(JDS_T6)};
    c[0] = P::mul(c[0], T(JDS_S2));
#pragma unroll
    for (int k = 1; k <= 3; ++k) {
        const int kc = 8 - k;
        T t1 = P::add(c[k], c[kc]);
        T t2 = P::sub(c[k], c[kc]);
        c[k] = P::add(P::mul(TW[k - 1], t2), P::mul(TW[kc - 1], t1));
        c[kc] = P::sub(P::mul(TW[k - 1], t1), P::mul(TW[kc - 1], t2));
    }
    c[4] = P::mul(c[4], T(2.0 * JDS_T3));
    T g[8];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        T x0 = c[k], x1 = c[k + 2], x2 = c[k + 4], x3 = c[k + 6];
        T tr1 = P::add(x3, x1);
        g[4 * k + 2] = P::sub(x3, x1);
        T tr2 = P::add(x0, x2);
        g[4 * k + 1] = P::sub(x0, x2);
        g[4 * k] = P::add(tr2, tr1);
        g[4 * k + 3] = P::sub(tr2, tr1);
    }
    T r[8];
    r[0] = P::add(g[0], g[4]);
    r[7] = P::sub(g[0], g[4]);
    r[4] = -g[7];
    r[3] = g[3];
    T tr = P::add(P::mul(T(JDS_WR), g[5]), P::mul(T(JDS_WI), g[6]));
    T ti = P::sub(P::mul(T(JDS_WR), g[6]), P::mul(T(JDS_WI), g[5]));
    r[1] = P::add(g[1], tr);
    r[5] = P::sub(g[1], tr);
    r[2] = P::add(ti, g[2]);
    r[6] = P::sub(ti, g[2]);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = P::mul(r[i], f);
#pragma unroll
    for (int k = 1; k <= 5; k += 2) {
        T a = r[k], b = r[k + 1];
        r[k] = P::sub(a, b);
        r[k + 1] = P::add(a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = r[i];
}

// ------------------------------------------------------------------------------
// The same transforms with every power-of-two multiplication DEFERRED (A5: "multiplications
// by 2, 0.5, 1/16 are exact and may be moved / merged").  A multiplication by 2^k is exact and
// commutes with every individually rounded operation - RN(2a + 2b) = 2 RN(a + b),
// RN(T (2a)) = 2 RN(T a) - so the doublings of dct8_ref (c0, c7, h3, h7, p, q), its f scale
// and the final halvings can be pulled out of the butterfly:
//   dct8_ref(c, f)[k]  ==  dct8_pow2_scale(k) * f * dct8_ref_unscaled(c)[k]     bit for bit,
// with dct8_pow2_scale(0) = dct8_pow2_scale(4) = 2 and 1 otherwise.  56 operations instead
// of 78 (the block codec folds the scales into its quantiser tables, BlockCodec<Exact>).
// ------------------------------------------------------------------------------
JDS_HD constexpr int dct8_pow2_shift(int k) { return (k == 0 || k == 4) ? 1 : 0; }   // log2 of the scale

template <class P>
JDS_HD void dct8_ref_unscaled(typename P::T* c) {
    typedef typename P::T T;
    const T TW[7] = {T(JDS_T0), T(JDS_T1), T(JDS_T2), T(JDS_T3), T(JDS_T4), T(JDS_T5), T(JDS_T6)};
    const T s07 = P::add(c[0], c[7]), d07 = P::sub(c[0], c[7]);       // h0 / 2, h4 / 2
#pragma unroll
    for (int k = 1; k <= 5; k += 2) {
        T a = c[k + 1], b = c[k];
        c[k + 1] = P::sub(a, b);
        c[k] = P::add(a, b);
    }
    const T h1 = P::add(c[1], c[5]);
    const T tr = P::sub(c[1], c[5]);
    const T ti = P::add(c[2], c[6]);
    const T h2 = P::sub(c[2], c[6]);
    const T h6 = P::add(P::mul(T(JDS_WR), ti), P::mul(T(JDS_WI), tr));
    const T h5 = P::sub(P::mul(T(JDS_WR), tr), P::mul(T(JDS_WI), ti));
    T r[8];                                                           // r / 2 of dct8_ref
    {
        const T u = P::add(s07, c[3]), v = P::sub(s07, c[3]);         // h3 = 2 c3
        r[0] = P::add(u, h1);
        r[4] = P::sub(u, h1);
        r[6] = P::add(v, h2);
        r[2] = P::sub(v, h2);
    }
    {
        const T u = P::sub(d07, c[4]), v = P::add(d07, c[4]);         // h7 = -2 c4
        r[1] = P::add(u, h5);
        r[5] = P::sub(u, h5);
        r[7] = P::add(v, h6);
        r[3] = P::sub(v, h6);
    }
    c[0] = P::mul(r[0], T(JDS_HALF_S2));
#pragma unroll
    for (int k = 1; k <= 3; ++k) {
        const int kc = 8 - k;
        T t1 = P::add(P::mul(TW[k - 1], r[kc]), P::mul(TW[kc - 1], r[k]));
        T t2 = P::sub(P::mul(TW[k - 1], r[k]), P::mul(TW[kc - 1], r[kc]));
        c[k] = P::add(t1, t2);
        c[kc] = P::sub(t1, t2);
    }
    c[4] = P::mul(r[4], TW[3]);
}

// idct8_ref without its f scale: idct8_ref(c, f) == idct8_ref_unscaled(f * c) bit for bit (f a
// power of two), so the caller scales the INPUT once (the block codec: in its dequantiser table)
template <class P>
JDS_HD void idct8_ref_unscaled(typename P::T* c) {
    typedef typename P::T T;
    const T TW[7] = {T(JDS_T0), T(JDS_T1), T(JDS_T2), T(JDS_T3), T(JDS_T4), T(JDS_T5), T(JDS_T6)};
    c[0] = P::mul(c[0], T(JDS_S2));
#pragma unroll
    for (int k = 1; k <= 3; ++k) {
        const int kc = 8 - k;
        T t1 = P::add(c[k], c[kc]);
        T t2 = P::sub(c[k], c[kc]);
        c[k] = P::add(P::mul(TW[k - 1], t2), P::mul(TW[kc - 1], t1));
        c[kc] = P::sub(P::mul(TW[k - 1], t1), P::mul(TW[kc - 1], t2));
    }
    c[4] = P::mul(c[4], T(2.0 * JDS_T3));
    T g[8];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        T x0 = c[k], x1 = c[k + 2], x2 = c[k + 4], x3 = c[k + 6];
        T tr1 = P::add(x3, x1);
        g[4 * k + 2] = P::sub(x3, x1);
        T tr2 = P::add(x0, x2);
        g[4 * k + 1] = P::sub(x0, x2);
        g[4 * k] = P::add(tr2, tr1);
        g[4 * k + 3] = P::sub(tr2, tr1);
    }
    T r[8];
    r[0] = P::add(g[0], g[4]);
    r[7] = P::sub(g[0], g[4]);
    r[4] = -g[7];
    r[3] = g[3];
    T tr = P::add(P::mul(T(JDS_WR), g[5]), P::mul(T(JDS_WI), g[6]));
    T ti = P::sub(P::mul(T(JDS_WR), g[6]), P::mul(T(JDS_WI), g[5]));
    r[1] = P::add(g[1], tr);
    r[5] = P::sub(g[1], tr);
    r[2] = P::add(ti, g[2]);
    r[6] = P::sub(ti, g[2]);
#pragma unroll
    for (int k = 1; k <= 5; k += 2) {
        T a = r[k], b = r[k + 1];
        r[k] = P::sub(a, b);
        r[k + 1] = P::add(a, b);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = r[i];
}

// ------------------------------------------------------------------------------
// 8-point transforms, fast: Arai-Agui-Nakajima scaled butterflies (5 multiplies).
// Forward output k equals the orthonormal DCT-II coefficient times
//   AAN_FWD[k] = 2*sqrt(2) * s_k,  s_0 = 1, s_k = sqrt(2) cos(k pi/16)   (per axis)
// and the inverse expects orthonormal coefficient k times
//   AAN_INV[k] = s_k / (2*sqrt(2))                                        (per axis)
// Both are folded into the quantiser tables (see jds_tables).
// ------------------------------------------------------------------------------
JDS_HD void dct8_aan(float* d) {
    float t0 = d[0] + d[7], t7 = d[0] - d[7];
    float t1 = d[1] + d[6], t6 = d[1] - d[6];
    float t2 = d[2] + d[5], t5 = d[2] - d[5];
    float t3 = d[3] + d[4], t4 = d[3] - d[4];
    float t10 = t0 + t3, t13 = t0 - t3;
    float t11 = t1 + t2, t12 = t1 - t2;
    d[0] = t10 + t11;
    d[4] = t10 - t11;
    float z1 = (t12 + t13) * 0.707106781186547524f;
    d[2] = t13 + z1;
    d[6] = t13 - z1;
    t10 = t4 + t5;
    t11 = t5 + t6;
    t12 = t6 + t7;
    float z5 = (t10 - t12) * 0.382683432365089772f;
    float z2 = fmaf(0.541196100146196985f, t10, z5);
    float z4 = fmaf(1.306562964876376527f, t12, z5);
    float z3 = t11 * 0.707106781186547524f;
    float z11 = t7 + z3, z13 = t7 - z3;
    d[5] = z13 + z2;
    d[3] = z13 - z2;
    d[1] = z11 + z4;
    d[7] = z11 - z4;
}

JDS_HD void idct8_aan(float* d) {
    float t10 = d[0] + d[4], t11 = d[0] - d[4];
    float t13 = d[2] + d[6];
    float t12 = fmaf(d[2] - d[6], 1.414213562373095049f, -t13);
    float t0 = t10 + t13, t3 = t10 - t13;
    float t1 = t11 + t12, t2 = t11 - t12;
    float z13 = d[5] + d[3], z10 = d[5] - d[3];
    float z11 = d[1] + d[7], z12 = d[1] - d[7];
    float t7 = z11 + z13;
    float t11b = (z11 - z13) * 1.414213562373095049f;
    float z5 = (z10 + z12) * 1.847759065022573512f;
    float t10b = fmaf(-1.082392200292393968f, z12, z5);   // = z5 - 1.0824 z12
    float t12b = fmaf(-2.613125929752753055f, z10, z5);   // = z5 - 2.6131 z10
    float t6 = t12b - t7;
    float t5 = t11b - t6;
    float t4 = t10b - t5;
    d[0] = t0 + t7;
    d[7] = t0 - t7;
    d[1] = t1 + t6;
    d[6] = t1 - t6;
    d[2] = t2 + t5;
    d[5] = t2 - t5;
    d[3] = t3 + t4;
    d[4] = t3 - t4;
}

// ------------------------------------------------------------------------------
// quantiser (A7) and the bit model (utils/metrics.py:75-79)
// ------------------------------------------------------------------------------
// bits of one non-zero coefficient: 6 position bits + ceil(log2(|v|+1)) + 1
JDS_HD int coeff_bits(int v) {
    int m = v < 0 ? -v : v;
    if (m == 0) return 0;
#if defined(__CUDA_ARCH__)
    return 7 + (32 - __clz(m));
#else
    int bl = 0;
    while (m) { ++bl; m >>= 1; }
    return 7 + bl;
#endif
}
// np.histogram(v, bins=50, range=(-100,100)): bin floor((v+100)/4), v=100 in bin 49
JDS_HD int hist_bin(int v) {
    if (v < -100 || v > 100) return -1;
    int b = (v + 100) >> 2;
    return b > 49 ? 49 : b;
}

// engines/quantizer.py:7-19 in plain fp64 (host side; integer-valued result)
inline void quant_table_host(int quality, double* table64) {
    static const double base[64] = {
        16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55,
        14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
        18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92,
        49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    volatile double scale = quality < 50 ? 5000.0 / (double)quality
                                         : 200.0 - 2.0 * (double)quality;
    for (int i = 0; i < 64; ++i) {
        volatile double prod = base[i] * scale;      // volatile: one rounding per op,
        volatile double sum = prod + 50.0;           // no contraction on any host
        volatile double quo = sum / 100.0;
        double q = floor(quo);
        if (q < 1.0) q = 1.0;
        if (q > 255.0) q = 255.0;
        table64[i] = q;
    }
}

}  // namespace jds
