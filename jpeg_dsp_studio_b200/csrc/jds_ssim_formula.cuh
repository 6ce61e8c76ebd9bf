// jds_ssim_formula.cuh - the per-window SSIM formula of k_ssim_strip (jds_ssim.cu), written on
// a packed-pair policy so that tests/emul can run the very same source on the CPU:
//   device: V = float2, every operation one packed f32x2 instruction (FADD2 / FMUL2 / FFMA2)
//   host:   V = two floats, the same operations with fmaf (individually rounded, like the GPU)
#pragma once
#include <math.h>
#include "jds_math.cuh"

#ifndef JDS_SSIM_COMPENSATED
#define JDS_SSIM_COMPENSATED 1
#endif

namespace jds {

#if defined(__CUDACC__)      // nvcc, both passes: the packed device flavour
struct Pair2 {
    typedef float2 V;
    static __device__ __forceinline__ V splat(float a) { return make_float2(a, a); }
    static __device__ __forceinline__ V add(V a, V b) { return __fadd2_rn(a, b); }
    static __device__ __forceinline__ V mul(V a, V b) { return __fmul2_rn(a, b); }
    static __device__ __forceinline__ V fma(V a, V b, V c) { return __ffma2_rn(a, b, c); }
    static __device__ __forceinline__ V rcp(V a) {              // MUFU.RCP, no range fix-up
        V r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(a.x));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(a.y));
        return r;
    }
};
#else                        // plain C++ (tests/emul)
struct HostPair {
    float x, y;
};
struct Pair2 {
    typedef HostPair V;
    static inline V splat(float a) { return V{a, a}; }
    static inline V add(V a, V b) { volatile float x = a.x + b.x, y = a.y + b.y; return V{x, y}; }
    static inline V mul(V a, V b) { volatile float x = a.x * b.x, y = a.y * b.y; return V{x, y}; }
    static inline V fma(V a, V b, V c) { return V{::fmaf(a.x, b.x, c.x), ::fmaf(a.y, b.y, c.y)}; }
    static inline V rcp(V a) { return V{1.0f / a.x, 1.0f / a.y}; }   // MUFU.RCP is within 1 ulp of this
};
#endif

// 0.5 * SSIM of one window for both channels of a pair, added to `ssum`, from the centred
// sums over its 49 samples:  S = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)),
// v = 49/48 (E[ab]-E[a]E[b]), written on the raw sums (N = 49, Ux = N ux):
//   S = 2 A1 A2 / (B1 B2),  A1 = 2 Ux Uy + C1 N^2,  B1 = Ux^2 + Uy^2 + C1 N^2,
//   A2 = N Sxy - Sx Sy + C2 N (N-1) / 2,  B2 = N Sq - Sx^2 - Sy^2 + C2 N (N-1)
// A2 and B2 are formed NEGATED (packed f32x2 has no operand negation, so every subtraction
// would cost an instruction): (-A2)(-B2)^-1 has the same value, the constants carry the signs,
// and the last product is fused into the accumulation.  Ux / Uy stay explicit: they are exact
// small integers on dark content, where an expanded form cancels catastrophically.
//
// Cancellation.  On flat content far from mid-grey Sx Sy ~ N Sxy ~ 4e7 while their difference
// is a few hundred: formed naively in fp32 (ulp 4 at 4e7) the covariance term is off by up to
// 5e-5 of C2 N (N-1) / 2 with a systematic sign, which showed as 1.6e-5 in the mean SSIM of a
// dark frame (tolerance 1e-5).  The compensated form below is exact for the integer channels:
//   p  = RN(Sx Sy),  e = Sx Sy - p            (one FMA: the exact rounding error)
//   N Sxy - p                                 (one FMA: the exact value is small, so no rounding)
//   A2 = (N Sxy - p) - e + C2 N (N-1) / 2
//   B2 = 2 A2 + D,   D = N Sum (x-y)^2 - (Sum (x-y))^2 = N (Sq - 2 Sxy) - (Sx - Sy)^2
// (vx + vy = 2 vxy + var(x - y); the constants match because C2 N (N-1) = 2 * C2 N (N-1) / 2);
// D involves only the DIFFERENCE of the images, which is small whenever SSIM matters.
// B1 reuses the squared difference: Ux^2 + Uy^2 = 2 Ux Uy + (Ux - Uy)^2.
// 19 packed instructions and two MUFU.RCP per window pair (14 for the naive form).
#if defined(__CUDACC__)
#define JDS_FORMULA_FN __device__ __forceinline__
#else
#define JDS_FORMULA_FN inline
#endif
template <class P>
JDS_FORMULA_FN typename P::V ssim_window_half_acc(typename P::V sx, typename P::V sy, typename P::V sq,
                                          typename P::V sc, typename P::V ssum) {
    typedef typename P::V V;
    constexpr float N = 49.0f;
    constexpr float C1N2 = 6.5025f * 2401.0f;
    constexpr float K2 = 58.5225f * 49.0f * 48.0f;
    const V Ux = P::add(sx, P::splat(128.0f * N)), Uy = P::add(sy, P::splat(128.0f * N));
    const V A1 = P::fma(P::mul(Ux, Uy), P::splat(2.0f), P::splat(C1N2));
#if JDS_SSIM_COMPENSATED
    const V nsx = P::mul(sx, P::splat(-1.0f));
    const V p = P::mul(sx, sy);
    const V en = P::fma(nsx, sy, p);                                            // p - Sx Sy, exact
    const V c1n = P::add(P::fma(sc, P::splat(-N), p), P::splat(-0.5f * K2));    // p - N Sxy - K2/2
    const V A2n = P::fma(en, P::splat(-1.0f), c1n);                             // -(N Sxy - Sx Sy + K2/2)
    const V sdn = P::add(nsx, sy);                                              // Sy - Sx = Uy - Ux
    const V sd2 = P::mul(sdn, sdn);
    const V B1 = P::add(A1, sd2);                      // Ux^2 + Uy^2 = 2 Ux Uy + (Ux - Uy)^2
    const V sdd = P::fma(sc, P::splat(-2.0f), sq);                              // Sum (x-y)^2
    const V Dn = P::fma(sdd, P::splat(-N), sd2);                                // -D
    const V B2n = P::fma(A2n, P::splat(2.0f), Dn);
#else
    const V B1 = P::fma(Ux, Ux, P::fma(Uy, Uy, P::splat(C1N2)));
    const V A2n = P::fma(sc, P::splat(-N), P::fma(sx, sy, P::splat(-0.5f * K2)));
    const V un = P::fma(sx, sx, P::fma(sy, sy, P::splat(-K2)));
    const V B2n = P::fma(sq, P::splat(-N), un);
#endif
    const V num = P::mul(A1, A2n), den = P::mul(B1, B2n);
    return P::fma(num, P::rcp(den), ssum);
}

}  // namespace jds
