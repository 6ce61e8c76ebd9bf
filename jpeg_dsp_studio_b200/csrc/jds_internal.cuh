// jds_internal.cuh - device-side structures shared by the kernels and the C ABI.
#pragma once
#include <stdint.h>
#include "jds_math.cuh"

namespace jds {

// Frame geometry.  Planes live in scratch with padded (multiple-of-8) strides.
struct Geom {
    int H, W;        // luma / image size
    int hc, wc;      // chroma plane size after decimation (engines/color_space.py:44-49)
    int Hp, Wp;      // luma padded to a multiple of 8 (engines/block_processor.py:7-16)
    int hcp, wcp;    // chroma padded
    int nbx_y, nby_y, nbx_c, nby_c;   // 8x8 blocks per plane
    int sub;         // JDS_SUB_*
    int W4;          // 4*floor(W/4): columns the OpenCV blur row pass does with FMA (A2)
    int general;     // chroma decimation by fractional-area weights (odd W, or odd H at 4:2:0)
    double sx, sy;   // wc/W, hc/H: cv2.resize(INTER_LINEAR) source step (A8)
    long long nblk_y, nblk_c;         // blocks per plane
    long long plane_y, plane_c;       // elements per padded plane
};

// Per-unit (frame or sweep point) accumulators, device resident, zeroed per call.
struct DevMetrics {
    unsigned long long sse_rgb;
    double sse_y;
    double ssim_sum[4];
    unsigned long long coeff_bits;
    unsigned long long nnz;
    unsigned long long hist[50];
};

// Power-of-two scale the exact block codec keeps OUT of forward coefficient i = r * 8 + c (r:
// vertical frequency, first axis, f = 1/16; c: horizontal frequency, second axis, f = 1):
// true coefficient = 2^-exact_coeff_shift(i) * unscaled coefficient (jds_math.cuh, dct8_ref_unscaled)
JDS_HD constexpr int exact_coeff_shift(int i) {
    return 4 - dct8_pow2_shift(i >> 3) - dct8_pow2_shift(i & 7);
}

// Quantiser tables of one unit, in the arithmetic the policy needs.
//   exact: q[i] = Q (fp64, integer valued), rq[i] = RN(1/Q), dqx[i] = Q * 2^-shift(i) / 16 (the
//          dequantiser with the deferred power-of-two scales of the transforms folded in,
//          BlockCodec<Exact>);   fq, dq unused
//   fast : fq[i] = 1 / (Q * AAN_FWD[u] * AAN_FWD[v]);  dq[i] = Q * AAN_INV[u] * AAN_INV[v]
struct QTables {
    double q[64];
    double rq[64];   // RN(1/Q): exact mode divides by Markstein's 3-operation sequence
    double dqx[64];  // Q * exact_coeff_scale(i) / 16, exact (powers of two)
    float fq[64];
    float dq[64];
};

}  // namespace jds
