// jds_ssim.cu - strip-streaming SSIM / SSE kernel (sm_100a).
//
// Computes, for R, G, B and BT.601 Y of two uint8 RGB images (original,
// reconstructed): the sum of the 7x7-window SSIM map over all windows inside the
// image (utils/metrics.py:12-14,21 -> skimage structural_similarity), and the sums
// of squared differences for PSNR (utils/metrics.py:11,20).
//
// Design (DESIGN.md "k_ssim_strip"):
//   * a CTA owns a vertical strip of 64 window columns (70 pixel columns) and walks
//     down it 7 rows at a time; nothing is recomputed vertically.
//   * rows arrive by TMA bulk copies (cp.async.bulk + mbarrier), double buffered.
//   * the four channels are processed as two PAIRS - (R,G) and (B,Y) - and every
//     floating-point operation is a packed f32x2 instruction (FADD2/FMUL2/FFMA2, new on
//     sm_100): one issue slot does the work for both channels of the pair.
//   * prep: bytes -> centred (x-128) fp32, channel pairs interleaved, plus BT.601 Y.
//   * pass 1 (horizontal): a thread owns (row, pair, 8-window segment) and forms the
//     7-tap window sums of x, y, x^2+y^2, xy by running prefix differences in
//     registers; results go to shared memory (32 B per row, pair, column).
//   * pass 2 (vertical): a thread owns (column, pair), keeps the last 7 horizontal
//     sums in a register ring, slides the 7-row sum and evaluates the SSIM formula.
//   * centring makes the fp32 sums of the integer channels exact and keeps the low
//     bits of the Y sums; the Y accumulators are rebuilt from the ring every chunk so
//     rounding cannot drift down a strip.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "jds_kernels.cuh"
#include "jds_ssim_formula.cuh"

// tuning knobs (defaults = the measured best; tools/ssim_variants.sh builds the others)
#ifndef JDS_SSIM_REBUILD_EVERY
#define JDS_SSIM_REBUILD_EVERY 1     // chunks between rebuilds of the non-integer (B,Y) accumulators
#endif
#ifndef JDS_SSIM_SPLIT_Q
#define JDS_SSIM_SPLIT_Q 0           // pass 1: separate x^2 / y^2 chains (more ILP, one more add per window)
#endif
#ifndef JDS_SSIM_CTAS_PER_SM
#define JDS_SSIM_CTAS_PER_SM 0       // 0: pick the vertical segmentation by the wave model below; n: aim at n CTAs per SM
#endif

namespace jds {

constexpr int S_OW = 64;             // window columns per strip
constexpr int S_LW = 80;             // pixel columns loaded per strip (multiple of 16)
constexpr int S_ROWB = S_LW * 3;     // bytes per loaded row
constexpr int S_R = 7;               // rows per chunk == window height
constexpr int S_NT = 128;            // threads per CTA = 64 columns x 2 channel pairs
constexpr int S_SEG = 8;             // windows per pass-1 task
constexpr int S_NSEG = S_OW / S_SEG; // 8

// Shared-memory rows are padded so that consecutive ROWS start 16 bytes apart modulo 128:
// a quarter-warp that walks 8 rows at the same column then hits 8 different bank groups.
constexpr int S_FHALF = S_LW / 4;      // float4 per half row of fpl (even / odd pixel pairs)
constexpr int S_FPITCH = 2 * S_FHALF + 1; // float4 row pitch of fpl: 81 * 16 B = 10 * 128 + 16
constexpr int S_HPITCH = 2 * S_OW + 1; // float4 row pitch of hxy / hqc: 129 * 16 B = 16 * 128 + 16

struct HSum {                        // window sums of one (row, pair, column)
    float2 sx, sy, sq, sc;           // .x = first channel of the pair, .y = second
};

struct SsimSmem {
    alignas(128) uint8_t raw[2][2][S_R][S_ROWB];   // [buffer][image][row][byte]
    // centred fp32 samples [image][pair][row]: each float4 holds two pixels x two channels
    // (c0 p, c1 p, c0 p+1, c1 p+1); even pixel-pairs in [0, 40), odd pixel-pairs in [40, 80)
    alignas(16) float4 fpl[2][2][S_R][S_FPITCH];
    alignas(16) float4 hxy[S_R][S_HPITCH];         // horizontal sums (sx, sy) [row][pair*64+col]
    alignas(16) float4 hqc[S_R][S_HPITCH];         // horizontal sums (sq, sc)
    alignas(8) unsigned long long bar[2];
    double red_ssim[4][2];
    double red_sse[4][4];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ int hswz(int col) { return col ^ ((col >> 3) & 7); }

// byte `b` (0..3) of `w` as (value - 128) in fp32: 0x4B0000vv is 2^23 + vv
template <int B>
__device__ __forceinline__ float byte_centered(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | B)) - 8388736.0f;
}

__device__ __forceinline__ float rcp_approx(float x) {        // MUFU.RCP, no range fix-up
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, f2(-1.0f), a); }

// the per-window formula lives in jds_ssim_formula.cuh (shared with the CPU emulation)
__device__ __forceinline__ float2 ssim_window_half2_acc(float2 sx, float2 sy, float2 sq, float2 sc,
                                                        float2 ssum) {
    return ssim_window_half_acc<Pair2>(sx, sy, sq, sc, ssum);
}

__global__ void __launch_bounds__(S_NT, 4)
k_ssim_strip(int H, int W, int seg_rows, const uint8_t* __restrict__ a_img, size_t a_stride,
             const uint8_t* __restrict__ b_img, size_t b_stride, DevMetrics* __restrict__ metrics,
             int want_ssim, int want_sse) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SsimSmem& sm = *reinterpret_cast<SsimSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int unit = blockIdx.z;
    const uint8_t* A = a_img + (size_t)unit * a_stride;
    const uint8_t* B = b_img + (size_t)unit * b_stride;
    const int x0 = blockIdx.x * S_OW;
    // rows: this CTA owns pixel rows [py0, py1) for the squared error and the window
    // rows [py0, min(py1, H-6)) for SSIM; it reads pixel rows [py0, min(py1 + 6, H))
    const int py0 = blockIdx.y * seg_rows;
    const int py1 = min(py0 + seg_rows, H);
    const int in_end = min(py1 + 6, H);
    const int n_rows = in_end - py0;
    const int n_chunks = (n_rows + S_R - 1) / S_R;
    const int load_px = min(S_LW, W - x0);          // multiple of 16
    const int own_px = min(S_OW, W - x0);           // pixels whose error this strip owns
    const uint32_t row_bytes = (uint32_t)load_px * 3u;
    const int nwin_x = W - 6 - x0;                  // window columns available from x0

    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int chunk) {
        // one warp issues the chunk's bulk copies (2 images x 7 rows), lane i -> row i
        const int buf = chunk & 1;
        const int y0 = py0 + chunk * S_R;
        const int nr = min(S_R, in_end - y0);
        if (tid == 0) mbar_expect_tx(&sm.bar[buf], row_bytes * 2u * (uint32_t)nr);
        __syncwarp();
        if (tid < 2 * S_R) {
            const int img = tid / S_R, r = tid % S_R;
            if (r < nr) {
                const uint8_t* src = (img ? B : A) + ((size_t)(y0 + r) * W + x0) * 3;
                bulk_g2s(&sm.raw[buf][img][r][0], src, row_bytes, &sm.bar[buf]);
            }
        }
    };
    if (tid < 32) {
        issue(0);
        if (n_chunks > 1) issue(1);
    }

    // pass-2 role: thread = (window column, channel pair)
    const int col = tid & (S_OW - 1);
    const int pair = tid >> 6;                      // warps 0,1: (R,G); warps 2,3: (B,Y)
    const int hidx = pair * S_OW + col;
    HSum ring[S_R];
#pragma unroll
    for (int i = 0; i < S_R; ++i) ring[i].sx = ring[i].sy = ring[i].sq = ring[i].sc = f2(0.f);
    HSum acc;
    acc.sx = acc.sy = acc.sq = acc.sc = f2(0.f);
    double ssim_a = 0.0, ssim_b = 0.0;
    const bool col_ok = want_ssim && col < nwin_x;
    // pass-1 role: thread = (row slot, segment, pair); slot 7 of every 8 idles, so a
    // quarter-warp is 7 rows of one segment: conflict free with the padded row pitches
    const int p1_row = tid & 7;
    const int p1_seg = (tid >> 3) & (S_NSEG - 1);
    const int p1_pair = tid >> 6;
    const bool p1_active = p1_row < S_R;
    const int p1_c0 = S_SEG * p1_seg;
    double sse_a = 0.0, sse_b = 0.0;                // squared error of the pair's channels

    // prep: bytes -> centred fp32, pairs (R,G) and (B,Y) interleaved; task = 4 pixels
    auto prep = [&](int chunk) {
        const int buf = chunk & 1;
        const int nrp = min(S_R, in_end - (py0 + chunk * S_R));
        mbar_wait(&sm.bar[buf], (uint32_t)((chunk >> 1) & 1));
        // only the 72 columns pass 1 reads (64 windows + 6, rounded to a group of 4) are
        // converted: 2 x 7 x 18 = 252 tasks = two rounds of the 128 threads
        constexpr int PG = (S_OW + 6 + 3) / 4;
        for (int task = tid; task < 2 * S_R * PG; task += S_NT) {
            const int img = task / (S_R * PG);
            const int rem = task % (S_R * PG);
            const int r = rem / PG, g4 = rem % PG;
            if (r < nrp) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(&sm.raw[buf][img][r][12 * g4]);
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                const float R0 = byte_centered<0>(w0), G0 = byte_centered<1>(w0), B0 = byte_centered<2>(w0);
                const float R1 = byte_centered<3>(w0), G1 = byte_centered<0>(w1), B1 = byte_centered<1>(w1);
                const float R2 = byte_centered<2>(w1), G2 = byte_centered<3>(w1), B2 = byte_centered<0>(w2);
                const float R3 = byte_centered<1>(w2), G3 = byte_centered<2>(w2), B3 = byte_centered<3>(w2);
                const float Y0 = fmaf(0.299f, R0, fmaf(0.587f, G0, 0.114f * B0));
                const float Y1 = fmaf(0.299f, R1, fmaf(0.587f, G1, 0.114f * B1));
                const float Y2 = fmaf(0.299f, R2, fmaf(0.587f, G2, 0.114f * B2));
                const float Y3 = fmaf(0.299f, R3, fmaf(0.587f, G3, 0.114f * B3));
                // pixels 4*g4 .. 4*g4+3 = pixel-pairs 2*g4 (even half) and 2*g4+1 (odd half):
                // consecutive threads store consecutive float4 -> no bank conflicts
                sm.fpl[img][0][r][g4] = make_float4(R0, G0, R1, G1);
                sm.fpl[img][0][r][S_FHALF + g4] = make_float4(R2, G2, R3, G3);
                sm.fpl[img][1][r][g4] = make_float4(B0, Y0, B1, Y1);
                sm.fpl[img][1][r][S_FHALF + g4] = make_float4(B2, Y2, B3, Y3);
            }
        }
    };
    prep(0);
    __syncthreads();

    // one vertical step of pass 2: take row r's horizontal sums, emit (optionally), drop the
    // row that leaves the 7-row window
#define JDS_P2_STEP(r, EMIT)                                                             \
    {                                                                                    \
        const float4 h0 = sm.hxy[r][hidx], h1 = sm.hqc[r][hidx];                         \
        HSum h;                                                                          \
        h.sx = f2(h0.x, h0.y); h.sy = f2(h0.z, h0.w);                                    \
        h.sq = f2(h1.x, h1.y); h.sc = f2(h1.z, h1.w);                                    \
        ring[r] = h;                                                                     \
        acc.sx = __fadd2_rn(acc.sx, h.sx);                                               \
        acc.sy = __fadd2_rn(acc.sy, h.sy);                                               \
        acc.sq = __fadd2_rn(acc.sq, h.sq);                                               \
        acc.sc = __fadd2_rn(acc.sc, h.sc);                                               \
        if (EMIT) ssum = ssim_window_half2_acc(acc.sx, acc.sy, acc.sq, acc.sc, ssum);       \
        const HSum o = ring[(r + 1) % S_R];                                              \
        acc.sx = sub2(acc.sx, o.sx);                                                     \
        acc.sy = sub2(acc.sy, o.sy);                                                     \
        acc.sq = sub2(acc.sq, o.sq);                                                     \
        acc.sc = sub2(acc.sc, o.sc);                                                     \
    }

    for (int c = 0; c < n_chunks; ++c) {
        // raw[c & 1] (chunk c) was consumed by prep(c) before the last barrier: refill it
        if (tid < 32 && c + 2 < n_chunks) issue(c + 2);
        const int y0 = py0 + c * S_R;
        const int nr = min(S_R, in_end - y0);

        // ---- pass 1: horizontal 7-tap sums of x, y, x^2+y^2, xy (sliding window) ---------
        if (p1_active && p1_row < nr) {
            // pixel-pair index of the segment's first pixel is 4 * seg (even): pairs alternate
            // between the even and the odd half of the row
            const float4* qa = &sm.fpl[0][p1_pair][p1_row][2 * p1_seg];
            const float4* qb = &sm.fpl[1][p1_pair][p1_row][2 * p1_seg];
            // sliding 7-tap window over 14 pixels, two pixels per 16-byte load; pixels are
            // loaded just before they enter the window and dropped 7 steps later, so only a
            // short history is live (register pressure decides the occupancy of this kernel)
            const bool own_row = (y0 + p1_row) < py1;
            const bool own_all = p1_c0 + S_SEG <= own_px;
            float2 xs[S_SEG + 6], ys[S_SEG + 6];
            float2 wx = f2(0.f), wy = f2(0.f), wq = f2(0.f), wc = f2(0.f), sse = f2(0.f);
#if JDS_SSIM_SPLIT_Q
            float2 wq2 = f2(0.f);
#endif
#pragma unroll
            for (int i2 = 0; i2 < (S_SEG + 6) / 2; ++i2) {
                const int off = (i2 & 1) * S_FHALF + (i2 >> 1);
                const float4 a = qa[off], b = qb[off];
                xs[2 * i2] = f2(a.x, a.y); xs[2 * i2 + 1] = f2(a.z, a.w);
                ys[2 * i2] = f2(b.x, b.y); ys[2 * i2 + 1] = f2(b.z, b.w);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int i = 2 * i2 + k;
                    if (i < S_SEG) {
                        // squared error of the 8 pixels this task owns (exact: integer channels)
                        const float2 d = sub2(xs[i], ys[i]);
                        if (own_all || p1_c0 + i < own_px) sse = __ffma2_rn(d, d, sse);
                    }
                    wx = __fadd2_rn(wx, xs[i]);
                    wy = __fadd2_rn(wy, ys[i]);
#if JDS_SSIM_SPLIT_Q
                    wq = __ffma2_rn(xs[i], xs[i], wq);
                    wq2 = __ffma2_rn(ys[i], ys[i], wq2);
#else
                    wq = __ffma2_rn(xs[i], xs[i], __ffma2_rn(ys[i], ys[i], wq));
#endif
                    wc = __ffma2_rn(xs[i], ys[i], wc);
                    if (i >= 6) {
                        const int j = i - 6;
                        sm.hxy[p1_row][p1_pair * S_OW + p1_c0 + j] = make_float4(wx.x, wx.y, wy.x, wy.y);
#if JDS_SSIM_SPLIT_Q
                        const float2 wqs = __fadd2_rn(wq, wq2);
                        sm.hqc[p1_row][p1_pair * S_OW + p1_c0 + j] = make_float4(wqs.x, wqs.y, wc.x, wc.y);
#else
                        sm.hqc[p1_row][p1_pair * S_OW + p1_c0 + j] = make_float4(wq.x, wq.y, wc.x, wc.y);
#endif
                        if (j < S_SEG - 1) {
                            const float2 ox = xs[j], oy = ys[j];
                            const float2 nox = sub2(f2(0.f), ox), noy = sub2(f2(0.f), oy);
                            wx = __fadd2_rn(wx, nox);
                            wy = __fadd2_rn(wy, noy);
#if JDS_SSIM_SPLIT_Q
                            wq = __ffma2_rn(nox, ox, wq);
                            wq2 = __ffma2_rn(noy, oy, wq2);
#else
                            wq = __ffma2_rn(nox, ox, __ffma2_rn(noy, oy, wq));
#endif
                            wc = __ffma2_rn(nox, oy, wc);
                        }
                    }
                }
            }
            if (own_row) {
                sse_a += (double)sse.x;
                sse_b += (double)sse.y;
            }
        }
        __syncthreads();

        // ---- pass 2: vertical sliding sum + SSIM; then prep of the next chunk -------------
        {
            if (pair == 1 && (JDS_SSIM_REBUILD_EVERY == 1 || (c % JDS_SSIM_REBUILD_EVERY) == 0)) {
                // rebuild the accumulators from the ring: the Y sums are not integers and
                // must not drift down the strip (warp-uniform branch)
                acc.sx = acc.sy = acc.sq = acc.sc = f2(0.f);
#pragma unroll
                for (int i = 1; i < S_R; ++i) {
                    acc.sx = __fadd2_rn(acc.sx, ring[i].sx);
                    acc.sy = __fadd2_rn(acc.sy, ring[i].sy);
                    acc.sq = __fadd2_rn(acc.sq, ring[i].sq);
                    acc.sc = __fadd2_rn(acc.sc, ring[i].sc);
                }
            }
            // rows r of this chunk in [e0, e1) complete a window row that this CTA owns
            const int e0 = max(0, 6 - c * S_R);
            const int e1 = want_ssim ? min(nr, py1 + 6 - y0) : 0;
            float2 ssum = f2(0.f);
            if (nr == S_R && e0 == 0 && e1 == S_R) {
                // steady state: all seven rows emit; straight-line code, no branches
                JDS_P2_STEP(0, true) JDS_P2_STEP(1, true) JDS_P2_STEP(2, true) JDS_P2_STEP(3, true)
                JDS_P2_STEP(4, true) JDS_P2_STEP(5, true) JDS_P2_STEP(6, true)
            } else {
#pragma unroll
                for (int r = 0; r < S_R; ++r)
                    if (r < nr) JDS_P2_STEP(r, (r >= e0 && r < e1))
            }
            if (col_ok) {
                ssim_a += (double)ssum.x;
                ssim_b += (double)ssum.y;
            }
        }
        if (c + 1 < n_chunks) prep(c + 1);
        __syncthreads();
    }
#undef JDS_P2_STEP

    // ---- reductions ---------------------------------------------------------------
    const int lane = tid & 31, warp = tid >> 5;
    // pass-1 pair varies inside a warp (56 tasks per pair): reduce per pair
    double e[4];
    e[0] = (p1_active && p1_pair == 0) ? sse_a : 0.0;
    e[1] = (p1_active && p1_pair == 0) ? sse_b : 0.0;
    e[2] = (p1_active && p1_pair == 1) ? sse_a : 0.0;
    e[3] = (p1_active && p1_pair == 1) ? sse_b : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ssim_a += __shfl_down_sync(0xffffffffu, ssim_a, o);
        ssim_b += __shfl_down_sync(0xffffffffu, ssim_b, o);
#pragma unroll
        for (int k = 0; k < 4; ++k) e[k] += __shfl_down_sync(0xffffffffu, e[k], o);
    }
    if (lane == 0) {
        sm.red_ssim[warp][0] = ssim_a;              // pass-2 pair = warp / 2
        sm.red_ssim[warp][1] = ssim_b;
#pragma unroll
        for (int k = 0; k < 4; ++k) sm.red_sse[warp][k] = e[k];
    }
    __syncthreads();
    if (tid < 4 && want_ssim) {
        // channel tid: pair = tid / 2 (warps 2*pair, 2*pair+1), half = tid % 2
        const int pr = tid >> 1, hf = tid & 1;
        atomicAdd(&metrics[unit].ssim_sum[tid],
                  2.0 * (sm.red_ssim[2 * pr][hf] + sm.red_ssim[2 * pr + 1][hf]));
    }
    if (tid == 32 && want_sse) {
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (int w = 0; w < 4; ++w)
            for (int k = 0; k < 4; ++k) t[k] += sm.red_sse[w][k];
        // integer channels: every partial is an exact integer below 2^53
        atomicAdd(&metrics[unit].sse_rgb, (unsigned long long)(t[0] + t[1] + t[2] + 0.5));
        atomicAdd(&metrics[unit].sse_y, t[3]);
    }
}

size_t ssim_strip_smem_bytes() { return sizeof(SsimSmem); }

// once per context (jds_ctx_create): dynamic shared memory opt-in and the largest carve-out so
// that four 54 KB CTAs fit an SM; idempotent, no state shared between contexts
cudaError_t ssim_configure_device() {
    cudaError_t e = cudaFuncSetAttribute(k_ssim_strip, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(SsimSmem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_ssim_strip, cudaFuncAttributePreferredSharedMemoryCarveout,
                                cudaSharedmemCarveoutMaxShared);
}

bool ssim_strip_supported(int H, int W, const void* a, size_t a_stride, const void* b,
                          size_t b_stride) {
    if (H < 7 || W < 7 || (W % 16) != 0) return false;
    if (((uintptr_t)a | (uintptr_t)b | a_stride | b_stride) & 15) return false;
    return true;
}

cudaError_t launch_ssim_strip(int H, int W, const uint8_t* a, size_t a_stride, const uint8_t* b,
                              size_t b_stride, DevMetrics* metrics, int units, bool want_ssim,
                              bool want_sse, int sm_count, cudaStream_t s) {
    const size_t smem = sizeof(SsimSmem);
    const int strips = (W + S_OW - 1) / S_OW;
    // Vertical segments per strip.  More segments = more CTAs to balance over the 4 x sm_count
    // resident slots, but every segment re-reads 6 rows and pays a fixed prologue; fewer = a
    // long ragged tail.  At least 126 rows per segment.
    int segs, seg_rows;
    auto rows_for = [&](int n) { return ((H + n - 1) / n + S_R - 1) / S_R * S_R; };
#if JDS_SSIM_CTAS_PER_SM > 0
    {
        const int want_ctas = sm_count * JDS_SSIM_CTAS_PER_SM;
        segs = (want_ctas + strips * units - 1) / (strips * units);
        if (segs < 1) segs = 1;
        seg_rows = rows_for(segs);
        if (seg_rows < 126) seg_rows = 126;
    }
#else
    {
        // time ~ work x (1 + c / r) / slots + (r + c) / 2  (r rows per segment, c ~ 28 row-times of
        // overlap + prologue per CTA, last term = average ragged tail) is flat around
        // r* = sqrt(2 c x work / slots); take the segment count just below it (measured on
        // 16 x 4K: 4 segments 1.40 ms, 2 segments 1.44 ms, 3 segments 1.43 ms)
        const double slots = 4.0 * sm_count;
        const double work = (double)strips * units * H;
        double r_opt = sqrt(2.0 * 28.0 * work / slots);
        if (r_opt < 126.0) r_opt = 126.0;
        segs = (int)((double)H / r_opt);
        if (segs < 1) segs = 1;
        seg_rows = rows_for(segs);
        if (seg_rows < 126) seg_rows = 126;
    }
#endif
    segs = (H + seg_rows - 1) / seg_rows;
    dim3 grid(strips, segs, units);
    k_ssim_strip<<<grid, S_NT, smem, s>>>(H, W, seg_rows, a, a_stride, b, b_stride, metrics,
                                          want_ssim ? 1 : 0, want_sse ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace jds
