// jds_fused.cu - fused fast-mode (fp32) codec kernels for sm_100a.
//
//   k_fast_chroma<SUB>   RGB tile -> decimated Cb/Cr (byte sums via IDP4A) -> 8x8 AAN DCT,
//                        quantise, bit model, dequantise, IDCT, clamp -> reconstructed
//                        chroma planes (fp32 in [0,1], L2 resident scratch)
//   k_fast_luma<SUB>     RGB tile -> Y -> 8x8 codec in registers -> (chroma tile from
//                        the planes above, bilinear upsample) -> YCbCr->RGB, clamp,
//                        truncate -> packed uint8 RGB
//   k_fast_444           all three channels of a tile in one CTA (no chroma planes)
//
// Common structure: one thread owns one 8x8 block (64 samples in registers, both DCT
// passes without any exchange); tiles are staged through shared memory with TMA bulk
// row copies (cp.async.bulk + mbarrier) so that global traffic is 128-byte coalesced;
// the block layout in shared memory has a 68-float stride so per-thread LDS.128 /
// STS.128 are conflict free.  Level shift, the 1/255 output scale and the AAN scale
// factors are folded into the tables / the DC term; the [0,255] clamps are the free
// .SAT modifier of the last butterfly add; uint8 truncation is one FFMA.RM against
// 2^23.  Squared errors and SSIM are left to k_ssim_strip (jds_ssim.cu), which reads
// the two uint8 images back through L2.
//
// Reference stages covered: engines/color_space.py:8-14,27-53,56-66,17-24;
// engines/dct_engine.py:17-27; engines/quantizer.py:22-29; utils/metrics.py:63-85
// (bit model); engines/pipeline.py:47-95.
#include <cuda_runtime.h>
#include <stdint.h>
#include "jds_kernels.cuh"
#include "jds_math.cuh"

namespace jds {

// ------------------------------------------------------------------------------
// TMA / mbarrier helpers
// ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void f_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void f_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(f_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void f_mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "F_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra F_DONE;\n"
        "bra F_WAIT_LOOP;\n"
        "F_DONE:\n"
        "}\n" ::"r"(f_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void f_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                           unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(f_smem_u32(dst)), "l"(src), "r"(bytes), "r"(f_smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------------------
// small arithmetic helpers
// ------------------------------------------------------------------------------
constexpr float kMagicRound = 12582912.0f;        // 1.5 * 2^23: round-half-even of |x| < 2^22
constexpr int kMagicRoundBits = 0x4B400000;
constexpr float kInv255 = 1.0f / 255.0f;
constexpr float k128_255 = 128.0f / 255.0f;
constexpr int BLK_STRIDE = 68;                    // floats per 8x8 block slot in shared memory
// luma tiles: which block rows keep their two 16-byte halves swapped - bit 3 of the slot index
// (the eight lanes of a quarter-warp that walks eight blocks of one row) XOR one bit of the row
// (SHIFT = 1 for 4:2:0, 0 for 4:2:2: a quarter-warp of the compose phase walks four blocks of two
// tasks whose rows differ in exactly that bit)
#ifndef JDS_LU_NEXT_PREFETCH
#define JDS_LU_NEXT_PREFETCH 592     // luma kernel: L2 prefetch distance in CTAs (0 = off); 592 = 4 CTAs x 148 SMs
#endif
#ifndef JDS_CA_PREFETCH
#define JDS_CA_PREFETCH 1
#endif
#ifndef JDS_LU_ROWSWZ
#define JDS_LU_ROWSWZ 1
#endif
template <int SHIFT>
__device__ __forceinline__ int slot_swz(int blk, int ry) {
    return JDS_LU_ROWSWZ ? (((blk >> 3) ^ (ry >> SHIFT)) & 1) : ((blk >> 3) & 1);
}

template <int B>
__device__ __forceinline__ float f_byte_centered(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | B)) - 8388736.0f;
}

// inverse AAN butterfly whose eight outputs are saturated to [0,1] (FADD.SAT)
__device__ __forceinline__ void idct8_aan_sat(float* d) {
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
    float t10b = fmaf(-1.082392200292393968f, z12, z5);
    float t12b = fmaf(-2.613125929752753055f, z10, z5);
    float t6 = t12b - t7;
    float t5 = t11b - t6;
    float t4 = t10b - t5;
    d[0] = __saturatef(t0 + t7);
    d[7] = __saturatef(t0 - t7);
    d[1] = __saturatef(t1 + t6);
    d[6] = __saturatef(t1 - t6);
    d[2] = __saturatef(t2 + t5);
    d[5] = __saturatef(t2 - t5);
    d[3] = __saturatef(t3 + t4);
    d[4] = __saturatef(t3 - t4);
}

// One 8x8 block, level-shifted samples in v (row-major) -> reconstructed samples / 255
// clamped to [0,1].  s_fq / s_dq: shared-memory tables (dq already divided by 255).
// esum accumulates the fp32 exponent fields of the non-zero quantised values
// (bits = esum - 119 * nnz, utils/metrics.py:75-79), nnz their count.
// Forward half of the block codec: level-shifted samples in v (row-major) -> scaled AAN DCT
// coefficients (quality independent: a sweep computes them once per frame).
__device__ __forceinline__ void codec_fast_fwd(float* v) {
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
}

// Back half: coefficients -> quantise, bit model, dequantise, IDCT -> reconstructed samples
// / 255 clamped to [0,1].  s_fq / s_dq: shared-memory tables (dq already divided by 255).
// esum accumulates the fp32 exponent fields of the non-zero quantised values
// (bits = esum - 119 * nnz, utils/metrics.py:75-79), nnz their count.
template <bool COEFFS>
__device__ __forceinline__ void codec_fast_back(float* v, const float* __restrict__ s_fq,
                                                const float* __restrict__ s_dq, unsigned& esum,
                                                unsigned& nnz, int16_t* __restrict__ coef_out) {
    unsigned e_acc = 0, n_acc = 0;
    uint32_t packed[32];
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
        const float4 fq = *reinterpret_cast<const float4*>(s_fq + i);
        const float4 dq = *reinterpret_cast<const float4*>(s_dq + i);
        const float fqv[4] = {fq.x, fq.y, fq.z, fq.w};
        const float dqv[4] = {dq.x, dq.y, dq.z, dq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float t = fmaf(v[i + k], fqv[k], kMagicRound);
            const float qf = t - kMagicRound;
            const unsigned e = (__float_as_uint(qf) >> 23) & 0xFFu;
            e_acc += e;
            n_acc += (e + 1u) >> 7;
            v[i + k] = qf * dqv[k];
            if (COEFFS) {
                const uint32_t qi = (uint32_t)(__float_as_int(t) - kMagicRoundBits) & 0xFFFFu;
                if (k & 1) packed[(i + k) >> 1] |= qi << 16;
                else packed[(i + k) >> 1] = qi;
            }
        }
    }
    esum += e_acc;
    nnz += n_acc;
    if (COEFFS) {
        uint4* out = reinterpret_cast<uint4*>(coef_out);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            out[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
    }
    v[0] += k128_255;                      // +128 (in 1/255 units) on every output sample
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
    for (int r = 0; r < 8; ++r) idct8_aan_sat(v + r * 8);
}

// One 8x8 block, level-shifted samples in v (row-major) -> reconstructed samples / 255
template <bool COEFFS>
__device__ __forceinline__ void codec_fast(float* v, const float* __restrict__ s_fq,
                                           const float* __restrict__ s_dq, unsigned& esum,
                                           unsigned& nnz, int16_t* __restrict__ coef_out) {
    codec_fast_fwd(v);
    codec_fast_back<COEFFS>(v, s_fq, s_dq, esum, nnz, coef_out);
}

// Hoisted sweeps (gui/worker.py:55-74 recomputes everything per quality; only quantisation
// onwards depends on it): the forward coefficients of a frame live in a scratch buffer, one
// 8 KB region per warp of the producing grid - float4 number i4 of the warp's lane L at
// (i4 * 32 + L) * 16 bytes, so every 16-byte access of a warp is one contiguous 512-byte piece.
// STAGE 0: the full kernel; 1: forward half only (writes the scratch, nothing else);
// 2: back half only (reads the scratch instead of the frame).
enum { STAGE_FULL = 0, STAGE_FWD = 1, STAGE_BACK = 2 };
constexpr size_t FCOEF_WARP_FLOATS = 32 * 64;

__device__ __forceinline__ float4* fcoef_warp_base(float* fcoef, int warps_per_cta) {
    const size_t cta = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    return reinterpret_cast<float4*>(fcoef + (cta * warps_per_cta + (threadIdx.x >> 5)) * FCOEF_WARP_FLOATS) +
           (threadIdx.x & 31);
}
__device__ __forceinline__ void fcoef_store(float4* base, const float* v) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
        base[i * 32] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void fcoef_load(const float4* base, float* v) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float4 a = base[i * 32];
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
}

__device__ __forceinline__ void load_tables(const QTables* __restrict__ src, float* s_fq,
                                            float* s_dq, int tid, int nthreads) {
    for (int i = tid; i < 64; i += nthreads) {
        s_fq[i] = src->fq[i];
        s_dq[i] = src->dq[i] * kInv255;
    }
}

__device__ __forceinline__ void flush_stats(unsigned esum, unsigned nnz, DevMetrics* m) {
    // bits = 6*nnz + sum(bit_length + 1) = esum - 119*nnz  (exponent = 126 + bit_length)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        esum += __shfl_down_sync(0xffffffffu, esum, o);
        nnz += __shfl_down_sync(0xffffffffu, nnz, o);
    }
    if ((threadIdx.x & 31) == 0 && nnz) {
        atomicAdd(&m->coeff_bits, (unsigned long long)esum - 119ull * nnz);
        atomicAdd(&m->nnz, (unsigned long long)nnz);
    }
}

// floor(255 * s) for s in [0,1] as the low byte of the returned word
__device__ __forceinline__ uint32_t trunc_byte(float s01) {
    return __float_as_uint(__fmaf_rd(s01, 255.0f, 8388608.0f));
}
__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// ------------------------------------------------------------------------------
// chroma kernel: SUB = 1 (4:2:2) or 2 (4:2:0), no prefilter
// tile = 16 x 4 chroma blocks per channel; thread t: channel t/64, block t%64
// ------------------------------------------------------------------------------
constexpr int CA_BX = 16, CA_BY = 4, CA_NT = 128;

template <int SUB>
struct ChromaSmem {
    static constexpr int ROWS = CA_BY * 8 * (SUB == 2 ? 2 : 1);
    alignas(16) float plane[2][CA_BX * CA_BY][BLK_STRIDE];
    alignas(16) float fq[64];
    alignas(16) float dq[64];
};

// codec of the chroma tile: thread t -> channel t/64, block t%64 of plane[2][64][68]
template <bool COEFFS, int STAGE>
__device__ __forceinline__ void chroma_codec_tail(
    const Geom& g, float (*plane)[CA_BX * CA_BY][BLK_STRIDE], const float* s_fq, const float* s_dq,
    int bx0, int by0, int unit, float* __restrict__ cplanes, size_t cplane_stride,
    int16_t* __restrict__ coeffs, size_t coeff_stride, DevMetrics* __restrict__ metrics,
    float* __restrict__ fcoef) {
    const int tid = threadIdx.x;
    const int ch = tid >> 6, blk = tid & 63;
    const int bx = bx0 + (blk & (CA_BX - 1)), by = by0 + (blk >> 4);
    unsigned esum = 0, nnz = 0;
    if (bx < g.nbx_c && by < g.nby_c) {
        float v[64];
        if (STAGE == STAGE_BACK) {
            fcoef_load(fcoef_warp_base(fcoef, CA_NT / 32), v);
        } else {
            const float4* src = reinterpret_cast<const float4*>(&plane[ch][blk][0]);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 a = src[i];
                v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
            }
            codec_fast_fwd(v);
        }
        if (STAGE == STAGE_FWD) {
            fcoef_store(fcoef_warp_base(fcoef, CA_NT / 32), v);
            return;
        }
        int16_t* cout = nullptr;
        if (COEFFS)
            cout = coeffs + (size_t)unit * coeff_stride +
                   ((size_t)g.nblk_y + (size_t)ch * g.nblk_c + (size_t)by * g.nbx_c + bx) * 64;
        codec_fast_back<COEFFS>(v, s_fq, s_dq, esum, nnz, cout);
        float* dst = cplanes + (size_t)unit * cplane_stride + (size_t)ch * g.plane_c +
                     (size_t)(by * 8) * g.wcp + bx * 8;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float4* d4 = reinterpret_cast<float4*>(dst + (size_t)r * g.wcp);
            d4[0] = make_float4(v[r * 8], v[r * 8 + 1], v[r * 8 + 2], v[r * 8 + 3]);
            d4[1] = make_float4(v[r * 8 + 4], v[r * 8 + 5], v[r * 8 + 6], v[r * 8 + 7]);
        }
    }
    if (STAGE != STAGE_FWD) flush_stats(esum, nnz, metrics + unit);
}

template <int SUB, bool COEFFS, int STAGE>
__global__ void __launch_bounds__(CA_NT)
k_fast_chroma(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
              float* __restrict__ cplanes, size_t cplane_stride,
              const QTables* __restrict__ tables, int table_stride,
              int16_t* __restrict__ coeffs, size_t coeff_stride, DevMetrics* __restrict__ metrics,
              float* __restrict__ fcoef) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ChromaSmem<SUB>& sm = *reinterpret_cast<ChromaSmem<SUB>*>(smem_raw);
    constexpr int VS = (SUB == 2) ? 2 : 1;            // luma rows per chroma row
    constexpr int ROWS = ChromaSmem<SUB>::ROWS;
    const int tid = threadIdx.x;
    const int unit = blockIdx.z;
    const uint8_t* in = rgb + (size_t)unit * rgb_stride;
    const int bx0 = blockIdx.x * CA_BX, by0 = blockIdx.y * CA_BY;   // chroma block origin
    const int x0 = bx0 * 16, y0 = by0 * 8 * VS;                      // luma pixel origin
    const int n_rows = min(ROWS, g.H - y0);
    const int n_px = min(CA_BX * 16, g.W - x0);

    if (STAGE != STAGE_FWD) load_tables(tables + (size_t)unit * table_stride, sm.fq, sm.dq, tid, CA_NT);
    if (STAGE == STAGE_BACK) {
        // hoisted sweep: the forward coefficients of this tile come from the scratch
        __syncthreads();
        chroma_codec_tail<COEFFS, STAGE>(g, sm.plane, sm.fq, sm.dq, bx0, by0, unit, cplanes, cplane_stride,
                                         coeffs, coeff_stride, metrics, fcoef);
        return;
    }

#if JDS_CA_PREFETCH
    // the tasks below fetch their rows two at a time: let L2 pull the whole tile in now
    {
        constexpr int LINES = CA_BX * 16 * 3 / 128;                   // 6 lines of 128 B per luma row
        const size_t row_end = (size_t)g.W * 3;
        for (int i = tid; i < ROWS * LINES; i += CA_NT) {
            const int r = i / LINES, l = i % LINES;
            const size_t off = (size_t)x0 * 3 + 128 * (size_t)l;
            if (r < n_rows && off < row_end)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(in + (size_t)(y0 + r) * row_end + off));
        }
    }
#endif
    // ---- decimation: a task = one chroma row x 8 chroma samples (16 luma pixels) ----
    // channel sums over the 2x1 / 2x2 footprint with IDP4A on the interleaved bytes,
    // accumulated on top of 2^23's bit pattern so the fp32 value is one FADD away
    constexpr int C_ROWS = CA_BY * 8, C_SEGS = CA_BX;
    const float ks = (SUB == 2) ? 0.25f : 0.5f;
    // thread t owns segment t%16 of chroma rows t/16 + 8k, k = 0..3; the loads of two tasks
    // are in flight at a time (software prefetch, 2 x VS x 3 sixteen-byte loads)
    constexpr int NTASK = C_ROWS * C_SEGS / CA_NT;                 // 4
    const int seg = tid & 15, crbase = tid >> 4;
    const bool seg_ok = seg * 16 < n_px;
    uint4 ld[2][VS][3];
    auto fetch = [&](int k, uint4 (&dst)[VS][3]) {
        const int cr = crbase + 8 * k;
        if (seg_ok && cr * VS < n_rows) {
#pragma unroll
            for (int v = 0; v < VS; ++v) {
                // 48 contiguous bytes per thread, consecutive threads contiguous: the three
                // 16-byte loads of a warp cover whole 128-byte lines between them
                const uint4* q = reinterpret_cast<const uint4*>(
                    in + ((size_t)(y0 + cr * VS + v) * g.W + x0 + seg * 16) * 3);
                dst[v][0] = __ldg(q);
                dst[v][1] = __ldg(q + 1);
                dst[v][2] = __ldg(q + 2);
            }
        }
    };
    fetch(0, ld[0]);
#pragma unroll
    for (int k = 0; k < NTASK; ++k) {
        if (k + 1 < NTASK) fetch(k + 1, ld[(k + 1) & 1]);
        const int cr = crbase + 8 * k;
        if (seg_ok && cr * VS < n_rows) {
            uint32_t w[VS][12];
#pragma unroll
            for (int v = 0; v < VS; ++v)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const uint4 a = ld[k & 1][v][i];
                    w[v][4 * i] = a.x; w[v][4 * i + 1] = a.y; w[v][4 * i + 2] = a.z; w[v][4 * i + 3] = a.w;
                }
            float cb[8], crr[8];
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {        // 12-byte group = 4 pixels = 2 chroma samples
                unsigned rA = 0x4B000000u, gA = 0x4B000000u, bA = 0x4B000000u;
                unsigned rB = 0x4B000000u, gB = 0x4B000000u, bB = 0x4B000000u;
#pragma unroll
                for (int v = 0; v < VS; ++v) {
                    const uint32_t w0 = w[v][3 * gq], w1 = w[v][3 * gq + 1], w2 = w[v][3 * gq + 2];
                    rA = __dp4a(w0, 0x01000001u, rA);                        // bytes 0,3
                    gA = __dp4a(w0, 0x00000100u, __dp4a(w1, 0x00000001u, gA)); // bytes 1,4
                    bA = __dp4a(w0, 0x00010000u, __dp4a(w1, 0x00000100u, bA)); // bytes 2,5
                    rB = __dp4a(w1, 0x00010000u, __dp4a(w2, 0x00000100u, rB)); // bytes 6,9
                    gB = __dp4a(w1, 0x01000000u, __dp4a(w2, 0x00010000u, gB)); // bytes 7,10
                    bB = __dp4a(w2, 0x01000001u, bB);                        // bytes 8,11
                }
                const float fr0 = __uint_as_float(rA) - 8388608.0f, fg0 = __uint_as_float(gA) - 8388608.0f,
                            fb0 = __uint_as_float(bA) - 8388608.0f;
                const float fr1 = __uint_as_float(rB) - 8388608.0f, fg1 = __uint_as_float(gB) - 8388608.0f,
                            fb1 = __uint_as_float(bB) - 8388608.0f;
                // level-shifted chroma (the +128 of colour_space.py:12-13 cancels the -128
                // of dct_engine.py:19)
                cb[2 * gq] = fmaf(-0.168736f * ks, fr0, fmaf(-0.331264f * ks, fg0, (0.5f * ks) * fb0));
                crr[2 * gq] = fmaf(0.5f * ks, fr0, fmaf(-0.418688f * ks, fg0, (-0.081312f * ks) * fb0));
                cb[2 * gq + 1] = fmaf(-0.168736f * ks, fr1, fmaf(-0.331264f * ks, fg1, (0.5f * ks) * fb1));
                crr[2 * gq + 1] = fmaf(0.5f * ks, fr1, fmaf(-0.418688f * ks, fg1, (-0.081312f * ks) * fb1));
            }
            const int blk = (cr >> 3) * CA_BX + seg, ry = cr & 7;
            float4* pb = reinterpret_cast<float4*>(&sm.plane[0][blk][ry * 8]);
            float4* pr = reinterpret_cast<float4*>(&sm.plane[1][blk][ry * 8]);
            pb[0] = make_float4(cb[0], cb[1], cb[2], cb[3]);
            pb[1] = make_float4(cb[4], cb[5], cb[6], cb[7]);
            pr[0] = make_float4(crr[0], crr[1], crr[2], crr[3]);
            pr[1] = make_float4(crr[4], crr[5], crr[6], crr[7]);
        }
    }
    __syncthreads();

    chroma_codec_tail<COEFFS, STAGE>(g, sm.plane, sm.fq, sm.dq, bx0, by0, unit, cplanes, cplane_stride,
                                     coeffs, coeff_stride, metrics, fcoef);
}

// ------------------------------------------------------------------------------
// chroma kernel with the anti-alias prefilter (use_prefilter=True):
// cv2.GaussianBlur(3x3, sigma 0.75, BORDER_REFLECT_101) of full-resolution Cb/Cr followed by
// the 2x1 / 2x2 INTER_AREA average (engines/color_space.py:38-49).  Blur and average are
// linear, so per axis they fold into one 4-tap filter [a, b, b, a], a = ke/2,
// b = (kc+ke)/2, on the pixel pair and its two neighbours (4:2:2 keeps the 3-tap blur
// vertically).  Phase A filters every luma row of the tile (+1 halo row each side)
// horizontally into shared memory, phase B combines rows.
// ------------------------------------------------------------------------------
// the tile's luma rows go through the row filter in this many passes (measured per 16 x 4K:
// 4:2:0 - 1 pass 0.442, 2 passes 0.370, 4 passes 0.346 ms; 4:2:2 - 0.509 / 0.464 / 0.495 ms)
template <int SUB>
struct ChromaPfSmem {
    static constexpr int PARTS = (SUB == 2) ? 4 : 2;
    static constexpr int LROWS = CA_BY * 8 * (SUB == 2 ? 2 : 1) + 2;     // luma rows incl. halo
    // rows of one pass (+ one halo row each side): the tile is filtered in PARTS passes so
    // that the row buffer stays small (4 passes: 18 KB instead of 68 KB -> four CTAs per SM
    // instead of two; the halo rows between passes are filtered twice: +12 %)
    static constexpr int PROWS = (LROWS - 2) / PARTS;
    // row-filtered (cb, cr) samples, level shifted: float4 number k2 (two samples) of segment `seg`
    // sits at [row][k2 * 16 + seg], so the 16 segments a half-warp works on are contiguous 16-byte
    // pieces (a [seg * 4 + k2] order puts them 64 B apart: 4-way bank conflicts, 67 % of all
    // shared-memory wavefronts of this kernel before the change)
    alignas(16) float4 hrow[PROWS + 2][4 * CA_BX];
    alignas(16) float plane[2][CA_BX * CA_BY][BLK_STRIDE];
    alignas(16) float fq[64];
    alignas(16) float dq[64];
};

template <int SUB, bool COEFFS, int STAGE>
__global__ void __launch_bounds__(CA_NT)
k_fast_chroma_pf(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
                 float* __restrict__ cplanes, size_t cplane_stride,
                 const QTables* __restrict__ tables, int table_stride,
                 int16_t* __restrict__ coeffs, size_t coeff_stride,
                 DevMetrics* __restrict__ metrics, float* __restrict__ fcoef) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ChromaPfSmem<SUB>& sm = *reinterpret_cast<ChromaPfSmem<SUB>*>(smem_raw);
    constexpr int VS = (SUB == 2) ? 2 : 1;
    constexpr int LROWS = ChromaPfSmem<SUB>::LROWS;
    constexpr float KE = (float)JDS_KE, KC = (float)JDS_KC;
    constexpr float HA = 0.5f * KE, HB = 0.5f * (KC + KE);                // folded 4-tap
    const int tid = threadIdx.x;
    const int unit = blockIdx.z;
    const uint8_t* in = rgb + (size_t)unit * rgb_stride;
    const int bx0 = blockIdx.x * CA_BX, by0 = blockIdx.y * CA_BY;
    const int x0 = bx0 * 16, y0 = by0 * 8 * VS;
    const int n_rows = min(LROWS - 2, g.H - y0);                          // tile's own luma rows
    const int n_px = min(CA_BX * 16, g.W - x0);
    if (STAGE != STAGE_FWD) load_tables(tables + (size_t)unit * table_stride, sm.fq, sm.dq, tid, CA_NT);

    constexpr int PROWS = ChromaPfSmem<SUB>::PROWS;
#if JDS_CA_PREFETCH
    // the row filter runs in passes: let L2 pull the rows of the later passes in while the first runs
    {
        constexpr int LINES = CA_BX * 16 * 3 / 128;
        const size_t row_end = (size_t)g.W * 3;
        for (int i = tid; i < (LROWS - 2) * LINES; i += CA_NT) {
            const int r = i / LINES, l = i % LINES;
            const size_t off = (size_t)x0 * 3 + 128 * (size_t)l;
            if (r >= PROWS && r < n_rows && off < row_end)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(in + (size_t)(y0 + r) * row_end + off));
        }
    }
#endif
    for (int part = 0; part < ChromaPfSmem<SUB>::PARTS; ++part) {
    const int lr0 = part * PROWS;                                         // first luma row (tile local, halo = -1)
    if (lr0 >= n_rows) break;
    if (part) __syncthreads();                                            // phase B of the last pass is done with hrow
    // ---- phase A: a task = one luma row (halo rows included) x 16 pixels -> 8 samples ----
    for (int task = tid; task < (PROWS + 2) * CA_BX; task += CA_NT) {
        const int lrp = task / CA_BX, seg = task % CA_BX;                 // row inside this pass
        const int lr = lr0 + lrp;                                         // row inside the tile (0 = halo above)
        if (lr >= n_rows + 2 || seg * 16 >= n_px) continue;
        // smem row lr holds image row y0 + lr - 1, REFLECT_101 at the top / bottom edge
        int y = y0 + lr - 1;
        y = y < 0 ? -y : (y >= g.H ? 2 * (g.H - 1) - y : y);
        const int xs = x0 + seg * 16;
        const uint8_t* row = in + ((size_t)y * g.W + xs) * 3;
        // 18 pixels xs-1 .. xs+16: the aligned 48 bytes plus the 16-byte chunks either side
        const bool has_l = xs > 0, has_r = xs + 16 < g.W;
        const uint4* q = reinterpret_cast<const uint4*>(row);
        const uint4 c0 = __ldg(q), c1 = __ldg(q + 1), c2 = __ldg(q + 2);
        const uint4 cl = has_l ? __ldg(q - 1) : make_uint4(0, 0, 0, 0);
        const uint4 cr4 = has_r ? __ldg(q + 3) : make_uint4(0, 0, 0, 0);
        const uint32_t w[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
        float cb[18], cr[18];                       // index i <-> pixel xs - 1 + i
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
            const uint32_t w0 = w[3 * gq], w1 = w[3 * gq + 1], w2 = w[3 * gq + 2];
            const float R[4] = {f_byte_centered<0>(w0), f_byte_centered<3>(w0), f_byte_centered<2>(w1), f_byte_centered<1>(w2)};
            const float G[4] = {f_byte_centered<1>(w0), f_byte_centered<0>(w1), f_byte_centered<3>(w1), f_byte_centered<2>(w2)};
            const float B[4] = {f_byte_centered<2>(w0), f_byte_centered<1>(w1), f_byte_centered<0>(w2), f_byte_centered<3>(w2)};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                cb[1 + 4 * gq + k] = fmaf(-0.168736f, R[k], fmaf(-0.331264f, G[k], 0.5f * B[k]));
                cr[1 + 4 * gq + k] = fmaf(0.5f, R[k], fmaf(-0.418688f, G[k], -0.081312f * B[k]));
            }
        }
        if (has_l) {                                // pixel xs-1 = bytes 13,14,15 of the chunk before
            const float r = f_byte_centered<1>(cl.w), gg = f_byte_centered<2>(cl.w), b = f_byte_centered<3>(cl.w);
            cb[0] = fmaf(-0.168736f, r, fmaf(-0.331264f, gg, 0.5f * b));
            cr[0] = fmaf(0.5f, r, fmaf(-0.418688f, gg, -0.081312f * b));
        } else {                                    // REFLECT_101: pixel -1 = pixel 1
            cb[0] = cb[2];
            cr[0] = cr[2];
        }
        if (has_r) {                                // pixel xs+16 = bytes 0,1,2 of the chunk after
            const float r = f_byte_centered<0>(cr4.x), gg = f_byte_centered<1>(cr4.x), b = f_byte_centered<2>(cr4.x);
            cb[17] = fmaf(-0.168736f, r, fmaf(-0.331264f, gg, 0.5f * b));
            cr[17] = fmaf(0.5f, r, fmaf(-0.418688f, gg, -0.081312f * b));
        } else {                                    // pixel W = pixel W-2
            cb[17] = cb[15];
            cr[17] = cr[15];
        }
        float4* dst = &sm.hrow[lrp][seg];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {            // two samples per float4 store
            float o[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = 2 * k2 + h;           // sample k: pixels 2k, 2k+1 -> cb[2k .. 2k+3]
                o[2 * h] = fmaf(HA, cb[2 * k] + cb[2 * k + 3], HB * (cb[2 * k + 1] + cb[2 * k + 2]));
                o[2 * h + 1] = fmaf(HA, cr[2 * k] + cr[2 * k + 3], HB * (cr[2 * k + 1] + cr[2 * k + 2]));
            }
            dst[k2 * CA_BX] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    __syncthreads();

    // ---- phase B: vertical taps, a task = one chroma row of this pass x 8 samples --------
    for (int task = tid; task < (PROWS / VS) * CA_BX; task += CA_NT) {
        const int crp = task / CA_BX, seg = task % CA_BX;                 // chroma row inside the pass
        const int cr_ = lr0 / VS + crp;                                   // chroma row inside the tile
        if (cr_ * VS >= n_rows || seg * 16 >= n_px) continue;
        float ob[8], orr[8];
        // smem rows of this chroma row: luma rows VS*cr_-1 .. VS*cr_+VS  ->  pass rows VS*crp .. +VS+1
        const float4* r0 = &sm.hrow[VS * crp][seg];
        const float4* r1 = &sm.hrow[VS * crp + 1][seg];
        const float4* r2 = &sm.hrow[VS * crp + 2][seg];
        const float4* r3 = &sm.hrow[VS * crp + (SUB == 2 ? 3 : 2)][seg];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            const float4 a = r0[k2 * CA_BX], b = r1[k2 * CA_BX], c = r2[k2 * CA_BX], d = r3[k2 * CA_BX];
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
            const float cv[4] = {c.x, c.y, c.z, c.w}, dv[4] = {d.x, d.y, d.z, d.w};
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                o[e] = (SUB == 2) ? fmaf(HA, av[e] + dv[e], HB * (bv[e] + cv[e]))
                                  : fmaf(KE, av[e] + cv[e], KC * bv[e]);
            ob[2 * k2] = o[0]; orr[2 * k2] = o[1];
            ob[2 * k2 + 1] = o[2]; orr[2 * k2 + 1] = o[3];
        }
        const int blk = (cr_ >> 3) * CA_BX + seg, ry = cr_ & 7;
        float4* pb = reinterpret_cast<float4*>(&sm.plane[0][blk][ry * 8]);
        float4* pr = reinterpret_cast<float4*>(&sm.plane[1][blk][ry * 8]);
        pb[0] = make_float4(ob[0], ob[1], ob[2], ob[3]);
        pb[1] = make_float4(ob[4], ob[5], ob[6], ob[7]);
        pr[0] = make_float4(orr[0], orr[1], orr[2], orr[3]);
        pr[1] = make_float4(orr[4], orr[5], orr[6], orr[7]);
    }
    }   // passes over the tile's rows
    __syncthreads();
    chroma_codec_tail<COEFFS, STAGE>(g, sm.plane, sm.fq, sm.dq, bx0, by0, unit, cplanes, cplane_stride,
                                     coeffs, coeff_stride, metrics, fcoef);
}

// ------------------------------------------------------------------------------
// luma + compose kernel: SUB = 1 (4:2:2) or 2 (4:2:0)
// tile = 32 x 4 luma blocks (256 x 32 pixels); thread t owns block t
// ------------------------------------------------------------------------------
constexpr int LU_BX = 32, LU_BY = 4, LU_NT = 128;
constexpr int LU_TW = LU_BX * 8, LU_TH = LU_BY * 8;      // 256 x 32
constexpr int CT_COLS = LU_TW / 2 + 8;                   // 136 chroma columns staged
// Row pitch of the staged chroma tile.  A compose warp reads two tile rows (lanes 0-15 / 16-31),
// lanes of a row 8 floats apart; 140 = 12 (mod 32) puts the two rows on disjoint bank sets, which
// halves the conflicts of the scalar loads of the outer samples (8-way -> 4-way; 136 = 8 (mod 32)
// mapped both rows onto the same four banks).  The 16-byte loads stay 2-way: inherent in lanes that
// sit 32 B apart (profiles/r2_luma_smem_conflicts.txt).
constexpr int CT_PITCH = CT_COLS + 4;

template <int SUB>
struct LumaSmem {
    static constexpr int CT_ROWS = (SUB == 2) ? LU_TH / 2 + 2 : LU_TH;
    alignas(16) float plane[LU_BX * LU_BY][BLK_STRIDE];
    alignas(128) float ctile[2][CT_ROWS][CT_PITCH];
    alignas(16) float fq[64];
    alignas(16) float dq[64];
    alignas(8) unsigned long long bar;
};

// 16 horizontally upsampled chroma samples from staged samples c[0..15], where c[4+k]
// is the chroma sample under output pair k (cv2.resize INTER_LINEAR 2x: weights .25/.75)
__device__ __forceinline__ void upsample_row16(const float* c, float* o) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float q = 0.75f * c[4 + k];
        o[2 * k] = fmaf(0.25f, c[3 + k], q);
        o[2 * k + 1] = fmaf(0.25f, c[5 + k], q);
    }
}

// load 16 staged chroma samples (k0-4 .. k0+11), replicate at the image border, upsample
__device__ __forceinline__ void stage_up16(const float* src, bool left_edge, bool right_edge,
                                           float* o) {
    float c[16];
    const float4* q = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 a = q[i];
        c[4 * i] = a.x; c[4 * i + 1] = a.y; c[4 * i + 2] = a.z; c[4 * i + 3] = a.w;
    }
    if (left_edge) c[3] = c[4];
    if (right_edge) c[12] = c[11];
    upsample_row16(c, o);
}

template <int SUB, bool COEFFS, int STAGE>
__global__ void __launch_bounds__(LU_NT)
k_fast_luma(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
            const float* __restrict__ cplanes, size_t cplane_stride,
            const QTables* __restrict__ tables, int table_stride,
            int16_t* __restrict__ coeffs, size_t coeff_stride,
            uint8_t* __restrict__ recon, size_t recon_stride, DevMetrics* __restrict__ metrics,
            float* __restrict__ fcoef) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    LumaSmem<SUB>& sm = *reinterpret_cast<LumaSmem<SUB>*>(smem_raw);
    const int tid = threadIdx.x;
    const int unit = blockIdx.z;
    const uint8_t* in = rgb + (size_t)unit * rgb_stride;
    const int x0 = blockIdx.x * LU_TW, y0 = blockIdx.y * LU_TH;
    const int n_rows = min(LU_TH, g.H - y0);
    const int n_px = min(LU_TW, g.W - x0);
    constexpr int SWZ_SHIFT = (SUB == 2) ? 1 : 0;             // row bit of the slot swizzle

    // chroma tile geometry: columns [ccol0, ccol0 + ncc), rows [crow0, crow0 + ncr)
    const int cx0 = x0 >> 1;
    const int cy0 = (SUB == 2) ? (y0 >> 1) : y0;
    const int ccol_lo = cx0 - 4;                              // smem column 0
    const int ccol0 = max(ccol_lo, 0);
    const int ccol1 = min(cx0 + (n_px >> 1) + 4, g.wc);
    const int crow_lo = (SUB == 2) ? cy0 - 1 : cy0;           // smem row 0
    const int crow0 = max(crow_lo, 0);
    const int crow1 = (SUB == 2) ? min(cy0 + (n_rows >> 1) + 1, g.hc) : min(cy0 + n_rows, g.hc);

    if (tid == 0) {
        f_mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (STAGE != STAGE_FWD && tid < 32) {
        // reconstructed chroma (written by k_fast_chroma earlier in this stream): TMA bulk
        // copies, one row per lane, overlapped with the luma work below
        if (tid == 0)
            f_mbar_expect_tx(&sm.bar, (uint32_t)(ccol1 - ccol0) * 4u * 2u * (uint32_t)(crow1 - crow0));
        __syncwarp();
        const float* cp = cplanes + (size_t)unit * cplane_stride;
        const int ncr = crow1 - crow0;
        for (int i = tid; i < 2 * ncr; i += 32) {
            const int chn = i / ncr, r = crow0 + i % ncr;
            f_bulk_g2s(&sm.ctile[chn][r - crow_lo][ccol0 - ccol_lo],
                       cp + (size_t)chn * g.plane_c + (size_t)r * g.wcp + ccol0,
                       (uint32_t)(ccol1 - ccol0) * 4u, &sm.bar);
        }
    }
    if (STAGE != STAGE_FWD) load_tables(tables + (size_t)unit * table_stride, sm.fq, sm.dq, tid, LU_NT);

    // ---- RGB -> level-shifted Y, block layout: a task = one row x 16 pixels ----------
    // thread t owns segment t%16 of rows t/16 + 8k, k = 0..3; all twelve 16-byte loads are
    // issued before the first use so the DRAM / L2 latency is paid once
    if (STAGE != STAGE_BACK) {
        constexpr int NTASK = LU_TH * (LU_TW / 16) / LU_NT;        // 4
        const int seg = tid & 15, rbase = tid >> 4;
        const bool seg_ok = seg * 16 < n_px;
        uint4 ld[NTASK][3];
#pragma unroll
        for (int k = 0; k < NTASK; ++k) {
            const int r = rbase + 8 * k;
            if (seg_ok && r < n_rows) {
                const uint4* q = reinterpret_cast<const uint4*>(
                    in + ((size_t)(y0 + r) * g.W + x0 + seg * 16) * 3);
                ld[k][0] = __ldg(q);
                ld[k][1] = __ldg(q + 1);
                ld[k][2] = __ldg(q + 2);
            }
        }
#if JDS_LU_NEXT_PREFETCH > 0
        {
            // Every CTA starts with DRAM-latency loads of its own tile (20 % of this kernel's stall
            // samples, directly and at the barrier behind them).  CTAs start in linear block order,
            // so this CTA asks L2 for the tile of the CTA that starts about one wave of resident
            // CTAs later - that CTA's first loads then hit L2.
            const int lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) + JDS_LU_NEXT_PREFETCH;
            const int tx = lin % gridDim.x, ty = (lin / gridDim.x) % gridDim.y, tz = lin / (gridDim.x * gridDim.y);
            if (tz < (int)gridDim.z) {
                const uint8_t* nin = rgb + (size_t)tz * rgb_stride;
                const int nx0 = tx * LU_TW, ny0 = ty * LU_TH;
                const size_t row_end = (size_t)g.W * 3;
                constexpr int LINES = LU_TW * 3 / 128;                 // 6 lines of 128 B per row
                for (int i = tid; i < LU_TH * LINES; i += LU_NT) {
                    const int r = i / LINES, l = i % LINES;
                    const size_t off = (size_t)nx0 * 3 + 128 * (size_t)l;
                    if (ny0 + r < g.H && off < row_end)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nin + (size_t)(ny0 + r) * row_end + off));
                }
            }
        }
#endif
#pragma unroll
        for (int k = 0; k < NTASK; ++k) {
            const int r = rbase + 8 * k;
            if (seg_ok && r < n_rows) {
                const uint32_t w[12] = {ld[k][0].x, ld[k][0].y, ld[k][0].z, ld[k][0].w,
                                        ld[k][1].x, ld[k][1].y, ld[k][1].z, ld[k][1].w,
                                        ld[k][2].x, ld[k][2].y, ld[k][2].z, ld[k][2].w};
                float yv[16];
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    const uint32_t w0 = w[3 * gq], w1 = w[3 * gq + 1], w2 = w[3 * gq + 2];
                    yv[4 * gq + 0] = fmaf(0.299f, f_byte_centered<0>(w0), fmaf(0.587f, f_byte_centered<1>(w0), 0.114f * f_byte_centered<2>(w0)));
                    yv[4 * gq + 1] = fmaf(0.299f, f_byte_centered<3>(w0), fmaf(0.587f, f_byte_centered<0>(w1), 0.114f * f_byte_centered<1>(w1)));
                    yv[4 * gq + 2] = fmaf(0.299f, f_byte_centered<2>(w1), fmaf(0.587f, f_byte_centered<3>(w1), 0.114f * f_byte_centered<0>(w2)));
                    yv[4 * gq + 3] = fmaf(0.299f, f_byte_centered<1>(w2), fmaf(0.587f, f_byte_centered<2>(w2), 0.114f * f_byte_centered<3>(w2)));
                }
                const int blk = (r >> 3) * LU_BX + seg * 2, ry = r & 7;
                // slot swizzle: the two 16-byte halves of a block row swap places in blocks with
                // bit 3 set (SLOT_SWZ), so the eight lanes of a quarter-warp - blocks 2s, s = 0..7,
                // 32 B apart modulo 128 - hit eight different bank groups instead of four twice
                const int f = slot_swz<SWZ_SHIFT>(blk, ry);
                float4* p0 = reinterpret_cast<float4*>(&sm.plane[blk][ry * 8]);
                float4* p1 = reinterpret_cast<float4*>(&sm.plane[blk + 1][ry * 8]);
                p0[f] = make_float4(yv[0], yv[1], yv[2], yv[3]);
                p0[1 - f] = make_float4(yv[4], yv[5], yv[6], yv[7]);
                p1[f] = make_float4(yv[8], yv[9], yv[10], yv[11]);
                p1[1 - f] = make_float4(yv[12], yv[13], yv[14], yv[15]);
            }
        }
    }
    __syncthreads();

    // ---- codec: one luma block per thread, in place -------------------------------
    {
        const int bx = (x0 >> 3) + (tid & (LU_BX - 1)), by = (y0 >> 3) + (tid >> 5);
        unsigned esum = 0, nnz = 0;
        if (bx < g.nbx_y && by < g.nby_y) {
            float v[64];
            float4* slot = reinterpret_cast<float4*>(&sm.plane[tid][0]);
            // swizzled halves (see above): float4 i = 2 * row + half lives at i ^ f(row), f = fb ^ g(row)
            // with fb fixed per block and g known at compile time - two bases, constant offsets:
            //   g = 0: even i at slot_e[i], odd i at slot_o[i];  g = 1: even i at slot_o[i+1], odd i at slot_e[i-1]
            const int f = slot_swz<SWZ_SHIFT>(tid, 0);
            float4* slot_e = slot + f;
            float4* slot_o = slot - f;
            auto at = [&](int i) -> float4* {
                const int g = JDS_LU_ROWSWZ ? (((i >> 1) >> SWZ_SHIFT) & 1) : 0;
                return g ? ((i & 1) ? slot_e + (i - 1) : slot_o + (i + 1)) : ((i & 1) ? slot_o + i : slot_e + i);
            };
            if (STAGE == STAGE_BACK) {
                fcoef_load(fcoef_warp_base(fcoef, LU_NT / 32), v);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float4 a = *at(i);
                    v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
                }
                codec_fast_fwd(v);
            }
            if (STAGE == STAGE_FWD) {
                fcoef_store(fcoef_warp_base(fcoef, LU_NT / 32), v);
                return;
            }
            int16_t* cout = nullptr;
            if (COEFFS)
                cout = coeffs + (size_t)unit * coeff_stride + ((size_t)by * g.nbx_y + bx) * 64;
            codec_fast_back<COEFFS>(v, sm.fq, sm.dq, esum, nnz, cout);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                *at(i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if (STAGE == STAGE_FWD) return;
        flush_stats(esum, nnz, metrics + unit);
    }
    __syncthreads();
    f_mbar_wait(&sm.bar, 0);

    // ---- compose: a task = 2 rows (SUB 2) or 1 row (SUB 1) x 16 pixels ---------------
    constexpr int RPT = (SUB == 2) ? 2 : 1;                   // rows per task
    uint8_t* out = recon + (size_t)unit * recon_stride;
    for (int task = tid; task < (LU_TH / RPT) * (LU_TW / 16); task += LU_NT) {
#if JDS_LU_ROWSWZ
        // a warp = 32 tasks = 16 segments x 2 consecutive rows / row pairs; every quarter-warp takes
        // FOUR segments of BOTH: its 16-byte loads of the chroma tile (rows 560 B = 48 mod 128 apart)
        // and of the Y slots (row-dependent swizzle) then cover all 32 banks once
        const int seg = (task & 3) + 4 * ((task >> 3) & 3), rp = 2 * (task >> 5) + ((task >> 2) & 1);
#else
        const int rp = task / (LU_TW / 16), seg = task % (LU_TW / 16);
#endif
        const int r0 = rp * RPT;
        if (r0 >= n_rows || seg * 16 >= n_px) continue;
        // staged chroma columns: smem column of chroma sample k0 is k0 - ccol_lo = 8*seg + 4
        const int sc = 8 * seg;                               // 16 floats from here: k0-4 .. k0+11
        const bool left_edge = (cx0 + 8 * seg) == 0;
        const bool right_edge = (cx0 + 8 * seg + 8) >= g.wc;
        float up[2][RPT][16];                                 // [channel][row][pixel]
#pragma unroll
        for (int chn = 0; chn < 2; ++chn) {
            if (SUB == 2) {
                const int cyg = cy0 + rp;                     // chroma row under this row pair
                const int ra = max(cyg - 1, 0) - crow_lo, rb = cyg - crow_lo,
                          rc = min(cyg + 1, g.hc - 1) - crow_lo;
                float ua[16], ub[16], uc[16];
                stage_up16(&sm.ctile[chn][ra][sc], left_edge, right_edge, ua);
                stage_up16(&sm.ctile[chn][rb][sc], left_edge, right_edge, ub);
                stage_up16(&sm.ctile[chn][rc][sc], left_edge, right_edge, uc);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float q = 0.75f * ub[i];
                    up[chn][0][i] = fmaf(0.25f, ua[i], q);    // even luma row: .25 above + .75 here
                    up[chn][RPT - 1][i] = fmaf(0.25f, uc[i], q);   // odd row: .75 here + .25 below
                }
            } else {
                stage_up16(&sm.ctile[chn][cy0 + r0 - crow_lo][sc], left_edge, right_edge, up[chn][0]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
            const int r = r0 + rr;
            if (r >= n_rows) break;
            const int blk = (r >> 3) * LU_BX + seg * 2, ry = r & 7;
            float yv[16];
            {
                const int f = slot_swz<SWZ_SHIFT>(blk, ry);
                const float4* p0 = reinterpret_cast<const float4*>(&sm.plane[blk][ry * 8]);
                const float4* p1 = reinterpret_cast<const float4*>(&sm.plane[blk + 1][ry * 8]);
                float4 a = p0[f], b = p0[1 - f], c = p1[f], d = p1[1 - f];
                yv[0] = a.x; yv[1] = a.y; yv[2] = a.z; yv[3] = a.w;
                yv[4] = b.x; yv[5] = b.y; yv[6] = b.z; yv[7] = b.w;
                yv[8] = c.x; yv[9] = c.y; yv[10] = c.z; yv[11] = c.w;
                yv[12] = d.x; yv[13] = d.y; yv[14] = d.z; yv[15] = d.w;
            }
            uint32_t bytes[48];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                // planes hold value/255 in [0,1]; chroma offset 128/255 folded into constants
                const float cbv = up[0][rr][i], crv = up[1][rr][i];
                const float rf = __saturatef(fmaf(1.402f, crv, yv[i] - 1.402f * k128_255));
                const float gf = __saturatef(fmaf(-0.344136f, cbv, fmaf(-0.714136f, crv, yv[i] + (0.344136f + 0.714136f) * k128_255)));
                const float bf = __saturatef(fmaf(1.772f, cbv, yv[i] - 1.772f * k128_255));
                bytes[3 * i] = trunc_byte(rf);
                bytes[3 * i + 1] = trunc_byte(gf);
                bytes[3 * i + 2] = trunc_byte(bf);
            }
            uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)(y0 + r) * g.W + x0 + seg * 16) * 3);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                uint4 o;
                o.x = pack4(bytes[16 * i + 0], bytes[16 * i + 1], bytes[16 * i + 2], bytes[16 * i + 3]);
                o.y = pack4(bytes[16 * i + 4], bytes[16 * i + 5], bytes[16 * i + 6], bytes[16 * i + 7]);
                o.z = pack4(bytes[16 * i + 8], bytes[16 * i + 9], bytes[16 * i + 10], bytes[16 * i + 11]);
                o.w = pack4(bytes[16 * i + 12], bytes[16 * i + 13], bytes[16 * i + 14], bytes[16 * i + 15]);
                dst[i] = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------
// 4:4:4 kernel: tile = 16 x 4 blocks (128 x 32 pixels); 192 threads = 64 blocks x 3
// ------------------------------------------------------------------------------
constexpr int F4_BX = 16, F4_BY = 4, F4_NT = 192;
constexpr int F4_TW = F4_BX * 8, F4_TH = F4_BY * 8;
constexpr int F4_ROWB = F4_TW * 3;                       // 384

struct F444Smem {
    alignas(128) uint8_t raw[F4_TH][F4_ROWB];
    alignas(16) float plane[3][F4_BX * F4_BY][BLK_STRIDE];
    alignas(16) float fq[64];
    alignas(16) float dq[64];
    alignas(8) unsigned long long bar;
};

template <bool COEFFS>
#ifndef JDS_F4_MIN_CTAS
#define JDS_F4_MIN_CTAS 3      // 65 KB of shared memory, 100 registers: three CTAs fit an SM once the carve-out is maximal
#endif
__global__ void __launch_bounds__(F4_NT, JDS_F4_MIN_CTAS)
k_fast_444(Geom g, const uint8_t* __restrict__ rgb, size_t rgb_stride,
           const QTables* __restrict__ tables, int table_stride,
           int16_t* __restrict__ coeffs, size_t coeff_stride,
           uint8_t* __restrict__ recon, size_t recon_stride, DevMetrics* __restrict__ metrics) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    F444Smem& sm = *reinterpret_cast<F444Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const int unit = blockIdx.z;
    const uint8_t* in = rgb + (size_t)unit * rgb_stride;
    const int x0 = blockIdx.x * F4_TW, y0 = blockIdx.y * F4_TH;
    const int n_rows = min(F4_TH, g.H - y0);
    const int n_px = min(F4_TW, g.W - x0);
    const uint32_t row_bytes = (uint32_t)n_px * 3u;
    if (tid == 0) {
        f_mbar_init(&sm.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) {
        if (tid == 0) f_mbar_expect_tx(&sm.bar, row_bytes * (uint32_t)n_rows);
        __syncwarp();
        for (int r = tid; r < n_rows; r += 32)
            f_bulk_g2s(&sm.raw[r][0], in + ((size_t)(y0 + r) * g.W + x0) * 3, row_bytes, &sm.bar);
    }
    load_tables(tables + (size_t)unit * table_stride, sm.fq, sm.dq, tid, F4_NT);
    f_mbar_wait(&sm.bar, 0);

    for (int task = tid; task < F4_TH * (F4_TW / 16); task += F4_NT) {
        const int r = task / (F4_TW / 16), seg = task % (F4_TW / 16);
        if (r < n_rows && seg * 16 < n_px) {
            const uint4* q = reinterpret_cast<const uint4*>(&sm.raw[r][seg * 48]);
            uint32_t w[12];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const uint4 a = q[i];
                w[4 * i] = a.x; w[4 * i + 1] = a.y; w[4 * i + 2] = a.z; w[4 * i + 3] = a.w;
            }
            float yv[16], cb[16], cr[16];
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
                const uint32_t w0 = w[3 * gq], w1 = w[3 * gq + 1], w2 = w[3 * gq + 2];
                const float R[4] = {f_byte_centered<0>(w0), f_byte_centered<3>(w0), f_byte_centered<2>(w1), f_byte_centered<1>(w2)};
                const float G[4] = {f_byte_centered<1>(w0), f_byte_centered<0>(w1), f_byte_centered<3>(w1), f_byte_centered<2>(w2)};
                const float B[4] = {f_byte_centered<2>(w0), f_byte_centered<1>(w1), f_byte_centered<0>(w2), f_byte_centered<3>(w2)};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // all three planes level-shifted by -128 (chroma: +128 then -128 cancel;
                    // the coefficient rows of Cb and Cr sum to zero so centring R,G,B is free)
                    yv[4 * gq + k] = fmaf(0.299f, R[k], fmaf(0.587f, G[k], 0.114f * B[k]));
                    cb[4 * gq + k] = fmaf(-0.168736f, R[k], fmaf(-0.331264f, G[k], 0.5f * B[k]));
                    cr[4 * gq + k] = fmaf(0.5f, R[k], fmaf(-0.418688f, G[k], -0.081312f * B[k]));
                }
            }
            const int blk = (r >> 3) * F4_BX + seg * 2, ry = r & 7;
            const float* srcs[3] = {yv, cb, cr};
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                float4* p0 = reinterpret_cast<float4*>(&sm.plane[p][blk][ry * 8]);
                float4* p1 = reinterpret_cast<float4*>(&sm.plane[p][blk + 1][ry * 8]);
                const float* s = srcs[p];
                p0[0] = make_float4(s[0], s[1], s[2], s[3]);
                p0[1] = make_float4(s[4], s[5], s[6], s[7]);
                p1[0] = make_float4(s[8], s[9], s[10], s[11]);
                p1[1] = make_float4(s[12], s[13], s[14], s[15]);
            }
        }
    }
    __syncthreads();
    {
        const int ch = tid >> 6, blk = tid & 63;
        const int bx = (x0 >> 3) + (blk & (F4_BX - 1)), by = (y0 >> 3) + (blk >> 4);
        unsigned esum = 0, nnz = 0;
        if (bx < g.nbx_y && by < g.nby_y) {
            float v[64];
            float4* slot = reinterpret_cast<float4*>(&sm.plane[ch][blk][0]);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float4 a = slot[i];
                v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
            }
            int16_t* cout = nullptr;
            if (COEFFS)
                cout = coeffs + (size_t)unit * coeff_stride +
                       ((size_t)ch * g.nblk_y + (size_t)by * g.nbx_y + bx) * 64;
            codec_fast<COEFFS>(v, sm.fq, sm.dq, esum, nnz, cout);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                slot[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        flush_stats(esum, nnz, metrics + unit);
    }
    __syncthreads();
    uint8_t* out = recon + (size_t)unit * recon_stride;
    for (int task = tid; task < F4_TH * (F4_TW / 16); task += F4_NT) {
        const int r = task / (F4_TW / 16), seg = task % (F4_TW / 16);
        if (r >= n_rows || seg * 16 >= n_px) continue;
        const int blk = (r >> 3) * F4_BX + seg * 2, ry = r & 7;
        float pl[3][16];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const float4* p0 = reinterpret_cast<const float4*>(&sm.plane[p][blk][ry * 8]);
            const float4* p1 = reinterpret_cast<const float4*>(&sm.plane[p][blk + 1][ry * 8]);
            float4 a = p0[0], b = p0[1], c = p1[0], d = p1[1];
            pl[p][0] = a.x; pl[p][1] = a.y; pl[p][2] = a.z; pl[p][3] = a.w;
            pl[p][4] = b.x; pl[p][5] = b.y; pl[p][6] = b.z; pl[p][7] = b.w;
            pl[p][8] = c.x; pl[p][9] = c.y; pl[p][10] = c.z; pl[p][11] = c.w;
            pl[p][12] = d.x; pl[p][13] = d.y; pl[p][14] = d.z; pl[p][15] = d.w;
        }
        uint32_t bytes[48];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float yv = pl[0][i], cbv = pl[1][i], crv = pl[2][i];
            const float rf = __saturatef(fmaf(1.402f, crv, yv - 1.402f * k128_255));
            const float gf = __saturatef(fmaf(-0.344136f, cbv, fmaf(-0.714136f, crv, yv + (0.344136f + 0.714136f) * k128_255)));
            const float bf = __saturatef(fmaf(1.772f, cbv, yv - 1.772f * k128_255));
            bytes[3 * i] = trunc_byte(rf);
            bytes[3 * i + 1] = trunc_byte(gf);
            bytes[3 * i + 2] = trunc_byte(bf);
        }
        uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)(y0 + r) * g.W + x0 + seg * 16) * 3);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            uint4 o;
            o.x = pack4(bytes[16 * i + 0], bytes[16 * i + 1], bytes[16 * i + 2], bytes[16 * i + 3]);
            o.y = pack4(bytes[16 * i + 4], bytes[16 * i + 5], bytes[16 * i + 6], bytes[16 * i + 7]);
            o.z = pack4(bytes[16 * i + 8], bytes[16 * i + 9], bytes[16 * i + 10], bytes[16 * i + 11]);
            o.w = pack4(bytes[16 * i + 12], bytes[16 * i + 13], bytes[16 * i + 14], bytes[16 * i + 15]);
            dst[i] = o;
        }
    }
}

// ------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------
bool fused_supported(const Geom& g, int prefilter, const void* rgb, size_t rgb_stride,
                     const void* recon, size_t recon_stride) {
    if (prefilter && g.sub != 0 && (g.H < 2 || g.W < 2)) return false;
    if ((g.W % 16) != 0 || (g.H % 8) != 0) return false;
    if (g.sub == 2 && (g.H % 16) != 0) return false;   // chroma planes must be whole blocks
    if (((uintptr_t)rgb | (uintptr_t)recon | rgb_stride | recon_stride) & 15) return false;
    return true;
}

size_t fused_chroma_plane_floats(const Geom& g) { return g.sub ? (size_t)2 * g.plane_c : 0; }

// Opt in to > 48 KB dynamic shared memory for every instantiation.  Called once per context
// from jds_ctx_create (the attribute is per device and the call is idempotent, so there is no
// shared mutable state between contexts / threads - the launch paths below set nothing).
cudaError_t fused_configure_device() {
    cudaError_t e;
#define JDS_SET(K, BYTES)                                                                       \
    if ((e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES))) != cudaSuccess) return e
    JDS_SET((k_fast_chroma<1, false, STAGE_FULL>), sizeof(ChromaSmem<1>));
    JDS_SET((k_fast_chroma<1, false, STAGE_FWD>), sizeof(ChromaSmem<1>));
    JDS_SET((k_fast_chroma<1, false, STAGE_BACK>), sizeof(ChromaSmem<1>));
    JDS_SET((k_fast_chroma<1, true, STAGE_FULL>), sizeof(ChromaSmem<1>));
    JDS_SET((k_fast_chroma<2, false, STAGE_FULL>), sizeof(ChromaSmem<2>));
    JDS_SET((k_fast_chroma<2, false, STAGE_FWD>), sizeof(ChromaSmem<2>));
    JDS_SET((k_fast_chroma<2, false, STAGE_BACK>), sizeof(ChromaSmem<2>));
    JDS_SET((k_fast_chroma<2, true, STAGE_FULL>), sizeof(ChromaSmem<2>));
    JDS_SET((k_fast_chroma_pf<1, false, STAGE_FULL>), sizeof(ChromaPfSmem<1>));
    JDS_SET((k_fast_chroma_pf<1, false, STAGE_FWD>), sizeof(ChromaPfSmem<1>));
    JDS_SET((k_fast_chroma_pf<1, true, STAGE_FULL>), sizeof(ChromaPfSmem<1>));
    JDS_SET((k_fast_chroma_pf<2, false, STAGE_FULL>), sizeof(ChromaPfSmem<2>));
    JDS_SET((k_fast_chroma_pf<2, false, STAGE_FWD>), sizeof(ChromaPfSmem<2>));
    JDS_SET((k_fast_chroma_pf<2, true, STAGE_FULL>), sizeof(ChromaPfSmem<2>));
    JDS_SET((k_fast_luma<1, false, STAGE_FULL>), sizeof(LumaSmem<1>));
    JDS_SET((k_fast_luma<1, false, STAGE_FWD>), sizeof(LumaSmem<1>));
    JDS_SET((k_fast_luma<1, false, STAGE_BACK>), sizeof(LumaSmem<1>));
    JDS_SET((k_fast_luma<1, true, STAGE_FULL>), sizeof(LumaSmem<1>));
    JDS_SET((k_fast_luma<2, false, STAGE_FULL>), sizeof(LumaSmem<2>));
    JDS_SET((k_fast_luma<2, false, STAGE_FWD>), sizeof(LumaSmem<2>));
    JDS_SET((k_fast_luma<2, false, STAGE_BACK>), sizeof(LumaSmem<2>));
    JDS_SET((k_fast_luma<2, true, STAGE_FULL>), sizeof(LumaSmem<2>));
    JDS_SET((k_fast_444<false>), sizeof(F444Smem));
    JDS_SET((k_fast_444<true>), sizeof(F444Smem));
    // the driver's default carve-out left room for two CTAs of this kernel only
    if ((e = cudaFuncSetAttribute(k_fast_444<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_fast_444<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
#undef JDS_SET
    return cudaSuccess;
}

// forward-coefficient scratch of a hoisted sweep: [chroma grid][luma grid], one 8 KB region per
// warp of the producing kernels
static size_t fcoef_chroma_floats(const Geom& g) {
    return g.sub == 0 ? 0 : (size_t)((g.nbx_c + CA_BX - 1) / CA_BX) * ((g.nby_c + CA_BY - 1) / CA_BY) *
                                (CA_NT / 32) * FCOEF_WARP_FLOATS;
}
size_t fused_fcoef_floats(const Geom& g) {
    if (g.sub == 0) return 0;                      // 4:4:4 sweeps are not hoisted
    return fcoef_chroma_floats(g) + (size_t)((g.nbx_y + LU_BX - 1) / LU_BX) * ((g.nby_y + LU_BY - 1) / LU_BY) *
                                        (LU_NT / 32) * FCOEF_WARP_FLOATS;
}

// stage: 0 full kernel; 1 forward half of ONE frame into `fcoef` (tables / planes / metrics
// unused); 2 back half of `units` quality points from `fcoef`
cudaError_t launch_fused_chroma(const Geom& g, int prefilter, const uint8_t* rgb, size_t rgb_stride,
                                float* cplanes, size_t cplane_stride, const QTables* tables,
                                int table_stride, int16_t* coeffs, size_t coeff_stride,
                                DevMetrics* metrics, int units, cudaStream_t s, int stage, float* fcoef) {
    dim3 grid((g.nbx_c + CA_BX - 1) / CA_BX, (g.nby_c + CA_BY - 1) / CA_BY, units);
#define JDS_ARGS g, rgb, rgb_stride, cplanes, cplane_stride, tables, table_stride, coeffs, coeff_stride, metrics, fcoef
    if (stage == STAGE_BACK) {
        // the coefficients already contain the prefilter: one kernel for both cases
        if (g.sub == 2) k_fast_chroma<2, false, STAGE_BACK><<<grid, CA_NT, sizeof(ChromaSmem<2>), s>>>(JDS_ARGS);
        else k_fast_chroma<1, false, STAGE_BACK><<<grid, CA_NT, sizeof(ChromaSmem<1>), s>>>(JDS_ARGS);
        return cudaGetLastError();
    }
    if (stage == STAGE_FWD) {
        if (prefilter) {
            if (g.sub == 2) k_fast_chroma_pf<2, false, STAGE_FWD><<<grid, CA_NT, sizeof(ChromaPfSmem<2>), s>>>(JDS_ARGS);
            else k_fast_chroma_pf<1, false, STAGE_FWD><<<grid, CA_NT, sizeof(ChromaPfSmem<1>), s>>>(JDS_ARGS);
        } else {
            if (g.sub == 2) k_fast_chroma<2, false, STAGE_FWD><<<grid, CA_NT, sizeof(ChromaSmem<2>), s>>>(JDS_ARGS);
            else k_fast_chroma<1, false, STAGE_FWD><<<grid, CA_NT, sizeof(ChromaSmem<1>), s>>>(JDS_ARGS);
        }
        return cudaGetLastError();
    }
    if (prefilter) {
#define JDS_LAUNCH_PF(SUBV, CO) \
    k_fast_chroma_pf<SUBV, CO, STAGE_FULL><<<grid, CA_NT, sizeof(ChromaPfSmem<SUBV>), s>>>(JDS_ARGS)
        if (g.sub == 2) { if (coeffs) JDS_LAUNCH_PF(2, true); else JDS_LAUNCH_PF(2, false); }
        else { if (coeffs) JDS_LAUNCH_PF(1, true); else JDS_LAUNCH_PF(1, false); }
#undef JDS_LAUNCH_PF
        return cudaGetLastError();
    }
#define JDS_LAUNCH_CA(SUBV, CO) \
    k_fast_chroma<SUBV, CO, STAGE_FULL><<<grid, CA_NT, sizeof(ChromaSmem<SUBV>), s>>>(JDS_ARGS)
    if (g.sub == 2) { if (coeffs) JDS_LAUNCH_CA(2, true); else JDS_LAUNCH_CA(2, false); }
    else { if (coeffs) JDS_LAUNCH_CA(1, true); else JDS_LAUNCH_CA(1, false); }
#undef JDS_LAUNCH_CA
#undef JDS_ARGS
    return cudaGetLastError();
}

cudaError_t launch_fused_luma(const Geom& g, const uint8_t* rgb, size_t rgb_stride,
                              const float* cplanes, size_t cplane_stride, const QTables* tables,
                              int table_stride, int16_t* coeffs, size_t coeff_stride,
                              uint8_t* recon, size_t recon_stride, DevMetrics* metrics, int units,
                              cudaStream_t s, int stage, float* fcoef) {
    if (g.sub == 0) {
        dim3 grid((g.nbx_y + F4_BX - 1) / F4_BX, (g.nby_y + F4_BY - 1) / F4_BY, units);
        if (coeffs) {
            k_fast_444<true><<<grid, F4_NT, sizeof(F444Smem), s>>>(
                g, rgb, rgb_stride, tables, table_stride, coeffs, coeff_stride, recon, recon_stride, metrics);
        } else {
            k_fast_444<false><<<grid, F4_NT, sizeof(F444Smem), s>>>(
                g, rgb, rgb_stride, tables, table_stride, coeffs, coeff_stride, recon, recon_stride, metrics);
        }
        return cudaGetLastError();
    }
    dim3 grid((g.nbx_y + LU_BX - 1) / LU_BX, (g.nby_y + LU_BY - 1) / LU_BY, units);
    float* lcoef = fcoef ? fcoef + fcoef_chroma_floats(g) : nullptr;
#define JDS_LAUNCH_LU(SUBV, CO, ST)                                                             \
    k_fast_luma<SUBV, CO, ST><<<grid, LU_NT, sizeof(LumaSmem<SUBV>), s>>>(                       \
        g, rgb, rgb_stride, cplanes, cplane_stride, tables, table_stride, coeffs,                \
        coeff_stride, recon, recon_stride, metrics, lcoef)
    if (stage == STAGE_FWD) {
        if (g.sub == 2) JDS_LAUNCH_LU(2, false, STAGE_FWD); else JDS_LAUNCH_LU(1, false, STAGE_FWD);
    } else if (stage == STAGE_BACK) {
        if (g.sub == 2) JDS_LAUNCH_LU(2, false, STAGE_BACK); else JDS_LAUNCH_LU(1, false, STAGE_BACK);
    } else if (g.sub == 2) {
        if (coeffs) JDS_LAUNCH_LU(2, true, STAGE_FULL); else JDS_LAUNCH_LU(2, false, STAGE_FULL);
    } else {
        if (coeffs) JDS_LAUNCH_LU(1, true, STAGE_FULL); else JDS_LAUNCH_LU(1, false, STAGE_FULL);
    }
#undef JDS_LAUNCH_LU
    return cudaGetLastError();
}

}  // namespace jds
