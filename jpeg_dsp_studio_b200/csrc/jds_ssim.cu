// jds_ssim.cu - strip-streaming SSIM / SSE kernel (sm_100a), second generation.
//
// Computes, for R, G, B and BT.601 Y of two uint8 RGB images (original,
// reconstructed): the sum of the 7x7-window SSIM map over all windows inside the
// image (utils/metrics.py:12-14,21 -> skimage structural_similarity), and the sums
// of squared differences for PSNR (utils/metrics.py:11,20).
//
// Design (DESIGN.md "k_ssim_strip"):
//   * a CTA owns a vertical strip of 64 window columns (70 pixel columns, 80 loaded) and
//     walks down it 7 rows at a time; nothing is recomputed vertically.
//   * rows arrive by TMA: ONE cp.async.bulk.tensor (3-D tensor map over bytes x rows x
//     units, box 240 B x 7 rows, SASS UTMALDG) per image and chunk, double buffered, issued
//     by one thread two chunks ahead; out-of-image rows / columns are zero-filled by the TMA
//     unit, so the kernel has no edge cases on the load side.
//   * the four channels are processed as two PAIRS - (R,G) and (B,Y) - and every
//     floating-point operation is a packed f32x2 instruction (FADD2/FMUL2/FFMA2, new on
//     sm_100): one issue slot does the work for both channels of the pair.
//   * pass 1 (horizontal): a thread owns (row, pair, 8-window segment) and forms the
//     7-tap window sums of x, y, x^2+y^2, xy by running differences in registers.  The
//     (R,G) threads read the RAW BYTES of their 14 pixels straight from the TMA tile (PRMT
//     into 2^23+v, one packed add centres both channels) - no staging pass, no float copy
//     of these two channels in shared memory; the (B,Y) threads read centred fp32 (B,Y)
//     pairs that `prep` derived from the bytes (BT.601 luma needs all three channels).
//   * pass 2 (vertical): a thread owns (column, pair), keeps the last 7 horizontal
//     sums in a register ring, slides the 7-row sum and evaluates the SSIM formula.
//   * prep of chunk c+1 is spread over the two barrier intervals of chunk c so that the
//     (R,G) warps (byte conversion in pass 1) and the (B,Y) warps carry the same load;
//     the (B,Y) planes are double buffered for that.
//   * centring makes the fp32 sums of the integer channels exact and keeps the low
//     bits of the Y sums; the Y accumulators are rebuilt from the ring every chunk so
//     rounding cannot drift down a strip.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include "jds_kernels.cuh"
#include "jds_ssim_formula.cuh"

// tuning knobs (defaults = the measured best; tools/ssim_variants.sh builds the others)
#ifndef JDS_SSIM_REBUILD_EVERY
#define JDS_SSIM_REBUILD_EVERY 2     // chunks between rebuilds of the non-integer (B,Y) accumulators (1: 1.208, 2: 1.199, 4: 1.195 ms)
#endif
#ifndef JDS_SSIM_CTAS_PER_SM
#define JDS_SSIM_CTAS_PER_SM 0       // 0: pick the vertical segmentation by the wave model below; n: aim at n CTAs per SM
#endif
#ifndef JDS_SSIM_ST64
#define JDS_SSIM_ST64 0              // 1: pass-1 results leave as 8-byte stores (no MOVs, twice the store instructions: 1.210 vs 1.208 ms)
#endif
#ifndef JDS_SSIM_ASM_STORE
#define JDS_SSIM_ASM_STORE 1         // pass-1 16-byte stores as inline PTX straight from the accumulators (1.195 ms; plain C++ stores: 1.241)
#endif
#ifndef JDS_SSIM_PREP_SPLIT
#define JDS_SSIM_PREP_SPLIT 0        // 1: spread prep(c+1) over both barrier intervals of chunk c
#endif
#ifndef JDS_SSIM_PAIRBAR
#define JDS_SSIM_PAIRBAR 0           // 1: pass 1 -> pass 2 hand-over by a 64-thread named barrier per channel pair (measured slower: 1.28 vs 1.21 ms)
#endif
#ifndef JDS_SSIM_PREP_BY1
#define JDS_SSIM_PREP_BY1 160        // with PAIRBAR: prep tasks [0, n) go to the (B,Y) threads, the rest to (R,G)
#endif
#ifndef JDS_SSIM_PREP_HOIST
#define JDS_SSIM_PREP_HOIST 1        // prep's per-task index arithmetic once per CTA instead of once per chunk
#endif
#ifndef JDS_SSIM_PREP_RG
#define JDS_SSIM_PREP_RG 0           // > 0: the (R,G) warps take only the first n of the 252 prep tasks, the (B,Y) warps the rest
#endif
#ifndef JDS_SSIM_SM_ALTERNATE
#define JDS_SSIM_SM_ALTERNATE 1      // successive CTAs of an SM swap which warps take (R,G) and which (B,Y)
#endif
#ifndef JDS_SSIM_SSE_FROM_SUMS
#define JDS_SSIM_SSE_FROM_SUMS 1     // squared error of a task's 8 pixels from its first window's sums + pixel 7
#endif

namespace jds {

constexpr int S_OW = 64;             // window columns per strip
constexpr int S_LW = 80;             // pixel columns loaded per strip (multiple of 16)
constexpr int S_ROWB = S_LW * 3;     // bytes per loaded row
constexpr int S_R = 7;               // rows per chunk == window height
constexpr int S_NT = 128;            // threads per CTA = 64 columns x 2 channel pairs
constexpr int S_SEG = 8;             // windows per pass-1 task
constexpr int S_NSEG = S_OW / S_SEG; // 8
constexpr int S_PG = (S_OW + 6 + 3) / 4;   // 18 groups of 4 pixels: the 72 columns pass 1 reads
constexpr int S_TILEB = 1792;        // bytes reserved per TMA tile (7 x 240 = 1680, rounded to 128)

// Shared-memory rows are padded so that consecutive ROWS start 16 bytes apart modulo 128:
// a quarter-warp that walks 8 rows at the same column then hits 8 different bank groups.
constexpr int S_BHALF = S_PG;              // float4 per half row of `by` (even / odd pixel pairs)
constexpr int S_BPITCH = 2 * S_BHALF + 1;  // 37 float4 = 592 B = 4 * 128 + 80
constexpr int S_HPITCH = 2 * S_OW + 1;     // float4 row pitch of hxy / hqc: 129 * 16 B = 16 * 128 + 16
// (B,Y) planes: double buffered only when prep(c+1) overlaps pass 1 of chunk c (PREP_SPLIT)
constexpr int S_NBY = JDS_SSIM_PREP_SPLIT ? 2 : 1;

struct HSum {                        // window sums of one (row, pair, column)
    float2 sx, sy, sq, sc;           // .x = first channel of the pair, .y = second
};

struct SsimSmem {
    alignas(128) uint8_t raw[2][2][S_TILEB];        // [buffer][image]: 7 rows x 240 B, as the TMA writes them
    // centred fp32 (B, Y) pairs [buffer][image][row]: each float4 holds two pixels
    // (B p, Y p, B p+1, Y p+1); even pixel-pairs in [0, 18), odd pixel-pairs in [18, 36)
    alignas(16) float4 by[S_NBY][2][S_R][S_BPITCH];
    alignas(16) float4 hxy[S_R][S_HPITCH];          // horizontal sums (sx, sy) [row][pair*64+col]
    alignas(16) float4 hqc[S_R][S_HPITCH];          // horizontal sums (sq, sc)
    alignas(8) unsigned long long bar[2];
    int swap;
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
// TMA tile load global -> shared through a tensor map, completion counted on an mbarrier
// (SASS: UTMALDG).  Coordinates: byte column, row, unit.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

// byte `B` (0..3) of `w` as 2^23 + value in fp32 bits (0x4B0000vv)
template <int B>
__device__ __forceinline__ float byte_biased(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | B));
}
// byte `B` (0..3) of `w` as (value - 128) in fp32
template <int B>
__device__ __forceinline__ float byte_centered(uint32_t w) {
    return byte_biased<B>(w) - 8388736.0f;
}

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, f2(-1.0f), a); }

// 8-byte shared-memory accesses straight from / into a packed register pair
__device__ __forceinline__ void st_shared_f2(uint32_t addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float2 ld_shared_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}

// the per-window formula lives in jds_ssim_formula.cuh (shared with the CPU emulation)
__device__ __forceinline__ float2 ssim_window_half2_acc(float2 sx, float2 sy, float2 sq, float2 sc,
                                                        float2 ssum) {
    return ssim_window_half_acc<Pair2>(sx, sy, sq, sc, ssum);
}

// (R, G) of pixel I of a 14-pixel segment held as 12 words of raw bytes, centred, as one pair
template <int I>
__device__ __forceinline__ float2 rg_centered(const uint32_t (&w)[12]) {
    constexpr int br = 3 * I, bg = 3 * I + 1;
    const float2 v = f2(byte_biased<(br & 3)>(w[br >> 2]), byte_biased<(bg & 3)>(w[bg >> 2]));
    return __fadd2_rn(v, f2(-8388736.0f));
}

// Pass 1 for one (row, segment) of channel pair PAIR: sliding 7-tap sums over 14 pixels ->
// 8 windows x (sx, sy, sq, sc) into hxy / hqc; returns the squared error of the 8 pixels the
// task owns (exact: integer channels / small sums).  Pixels are converted (PAIR 0: raw bytes
// of the TMA tile) or loaded (PAIR 1: the prepared (B,Y) plane) just before they enter the
// window and dropped 7 steps later, so only a short history is live.
template <int PAIR>
__device__ __forceinline__ float2 pass1_task(SsimSmem& sm, int buf, int p1_row, int p1_seg) {
    float2 xs[S_SEG + 6], ys[S_SEG + 6];
    uint32_t wa[12], wb[12];
    const float4 *qa = nullptr, *qb = nullptr;
    if (PAIR == 0) {
        // raw bytes of pixels 8*seg .. 8*seg+13 of this row: 42 bytes from byte 24*seg,
        // six 8-byte loads per image (a half-warp = 2 segments x 8 rows: conflict free)
        const uint2* pa = reinterpret_cast<const uint2*>(&sm.raw[buf][0][p1_row * S_ROWB + 24 * p1_seg]);
        const uint2* pb = reinterpret_cast<const uint2*>(&sm.raw[buf][1][p1_row * S_ROWB + 24 * p1_seg]);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const uint2 a = pa[k], b = pb[k];
            wa[2 * k] = a.x; wa[2 * k + 1] = a.y;
            wb[2 * k] = b.x; wb[2 * k + 1] = b.y;
        }
    } else {
        // pixel-pair index of the segment's first pixel is 4 * seg (even): pairs alternate
        // between the even and the odd half of the row
        qa = &sm.by[buf & (S_NBY - 1)][0][p1_row][2 * p1_seg];
        qb = &sm.by[buf & (S_NBY - 1)][1][p1_row][2 * p1_seg];
    }
    float2 wx = f2(0.f), wy = f2(0.f), wq = f2(0.f), wc = f2(0.f), sse = f2(0.f);
    const uint32_t hx_dst = smem_u32(&sm.hxy[p1_row][PAIR * S_OW + S_SEG * p1_seg]);
    const uint32_t hq_dst = smem_u32(&sm.hqc[p1_row][PAIR * S_OW + S_SEG * p1_seg]);
#if JDS_SSIM_ST64
    const uint32_t st_swap = (p1_seg & 1) * 8;
#define P1_STORE(j)                                                                      \
            st_shared_f2(hx_dst + 16 * (j) + st_swap, wx);                               \
            st_shared_f2(hx_dst + 16 * (j) + (8 - st_swap), wy);                         \
            st_shared_f2(hq_dst + 16 * (j) + st_swap, wq);                               \
            st_shared_f2(hq_dst + 16 * (j) + (8 - st_swap), wc);
#elif JDS_SSIM_ASM_STORE
#define P1_STORE(j)                                                                      \
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hx_dst + 16 * (j)), \
                         "f"(wx.x), "f"(wx.y), "f"(wy.x), "f"(wy.y) : "memory");         \
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(hq_dst + 16 * (j)), \
                         "f"(wq.x), "f"(wq.y), "f"(wc.x), "f"(wc.y) : "memory");
#else
    float4* const hx_p = &sm.hxy[p1_row][PAIR * S_OW + S_SEG * p1_seg];
    float4* const hq_p = &sm.hqc[p1_row][PAIR * S_OW + S_SEG * p1_seg];
#define P1_STORE(j)                                                                      \
            hx_p[j] = make_float4(wx.x, wx.y, wy.x, wy.y);                               \
            hq_p[j] = make_float4(wq.x, wq.y, wc.x, wc.y);
#endif
#define JDS_P1_PIXEL(I)                                                                  \
    {                                                                                    \
        constexpr int i = (I);                                                           \
        if (PAIR == 0) {                                                                 \
            xs[i] = rg_centered<i>(wa);                                                  \
            ys[i] = rg_centered<i>(wb);                                                  \
        } else if ((i & 1) == 0) {                                                       \
            constexpr int i2 = i >> 1;                                                   \
            constexpr int off = (i2 & 1) * S_BHALF + (i2 >> 1);                          \
            const float4 a = qa[off], b = qb[off];                                       \
            xs[i] = f2(a.x, a.y); xs[i + 1] = f2(a.z, a.w);                              \
            ys[i] = f2(b.x, b.y); ys[i + 1] = f2(b.z, b.w);                              \
        }                                                                                \
        if (JDS_SSIM_SSE_FROM_SUMS ? (i == S_SEG - 1) : (i < S_SEG)) {                   \
            const float2 d = sub2(xs[i], ys[i]);                                         \
            sse = __ffma2_rn(d, d, sse);                                                 \
        }                                                                                \
        wx = __fadd2_rn(wx, xs[i]);                                                      \
        wy = __fadd2_rn(wy, ys[i]);                                                      \
        wq = __ffma2_rn(xs[i], xs[i], __ffma2_rn(ys[i], ys[i], wq));                     \
        wc = __ffma2_rn(xs[i], ys[i], wc);                                               \
        if (i >= 6) {                                                                    \
            constexpr int j = i >= 6 ? i - 6 : 0;                                        \
            /* pixels 0..6 are window 0: sum (x-y)^2 = sq - 2 sc; pixel 7 was added above */ \
            if (JDS_SSIM_SSE_FROM_SUMS && j == 0) sse = __fadd2_rn(sse, __ffma2_rn(wc, f2(-2.0f), wq)); \
            P1_STORE(j)                                                                  \
            if (j < S_SEG - 1) {                                                         \
                const float2 ox = xs[j], oy = ys[j];                                     \
                const float2 nox = sub2(f2(0.f), ox), noy = sub2(f2(0.f), oy);           \
                wx = __fadd2_rn(wx, nox);                                                \
                wy = __fadd2_rn(wy, noy);                                                \
                wq = __ffma2_rn(nox, ox, __ffma2_rn(noy, oy, wq));                       \
                wc = __ffma2_rn(nox, oy, wc);                                            \
            }                                                                            \
        }                                                                                \
    }
    JDS_P1_PIXEL(0) JDS_P1_PIXEL(1) JDS_P1_PIXEL(2) JDS_P1_PIXEL(3) JDS_P1_PIXEL(4)
    JDS_P1_PIXEL(5) JDS_P1_PIXEL(6) JDS_P1_PIXEL(7) JDS_P1_PIXEL(8) JDS_P1_PIXEL(9)
    JDS_P1_PIXEL(10) JDS_P1_PIXEL(11) JDS_P1_PIXEL(12) JDS_P1_PIXEL(13)
#undef P1_STORE
#undef JDS_P1_PIXEL
    return sse;
}

#ifndef JDS_SSIM_MIN_CTAS
#define JDS_SSIM_MIN_CTAS 4
#endif
#if JDS_SSIM_SM_ALTERNATE
// The (R,G) warps convert bytes inside pass 1 and carry ~15 % more instructions than the (B,Y)
// warps.  Warps w of all resident CTAs share scheduler w % 4, so with a fixed assignment two of an
// SM's four schedulers would always hold the heavy warps.  Every CTA takes a turn number from its
// SM and odd turns swap the roles: each scheduler then sees both kinds.  Only a balance hint -
// results do not depend on it.
__device__ unsigned int g_ssim_sm_turn[1024];
#endif
__global__ void __launch_bounds__(S_NT, JDS_SSIM_MIN_CTAS)
k_ssim_strip(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             int H, int W, int seg_rows, int a_unit_step, int b_unit_step,
             DevMetrics* __restrict__ metrics, int want_ssim, int want_sse) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SsimSmem& sm = *reinterpret_cast<SsimSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int unit = blockIdx.z;
    const int x0 = blockIdx.x * S_OW;
    // rows: this CTA owns pixel rows [py0, py1) for the squared error and the window
    // rows [py0, min(py1, H-6)) for SSIM; it reads pixel rows [py0, min(py1 + 6, H))
    const int py0 = blockIdx.y * seg_rows;
    const int py1 = min(py0 + seg_rows, H);
    const int in_end = min(py1 + 6, H);
    const int n_rows = in_end - py0;
    const int n_chunks = (n_rows + S_R - 1) / S_R;
    const int own_px = min(S_OW, W - x0);           // pixels whose error this strip owns (multiple of 16)
    const int nwin_x = W - 6 - x0;                  // window columns available from x0

    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#if JDS_SSIM_SM_ALTERNATE
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        sm.swap = (int)(atomicAdd(&g_ssim_sm_turn[smid & 1023], 1u) & 1u);
#endif
    }
    __syncthreads();
#if JDS_SSIM_SM_ALTERNATE
    const int swap = sm.swap;
#else
    const int swap = 0;
#endif

    auto issue = [&](int chunk) {
        // one thread: two tile loads (original, reconstruction) of 7 rows x 240 bytes; rows past
        // the image and columns past its right edge arrive as zeros
        const int buf = chunk & 1;
        const int y0 = py0 + chunk * S_R;
        mbar_expect_tx(&sm.bar[buf], 2u * S_R * S_ROWB);
        tma_load_3d(&sm.raw[buf][0][0], &map_a, x0 * 3, y0, unit * a_unit_step, &sm.bar[buf]);
        tma_load_3d(&sm.raw[buf][1][0], &map_b, x0 * 3, y0, unit * b_unit_step, &sm.bar[buf]);
    };
    if (tid == 0) {
        issue(0);
        if (n_chunks > 1) issue(1);
    }

    // pass-2 role: thread = (window column, channel pair)
    const int col = tid & (S_OW - 1);
    const int pair = (tid >> 6) ^ swap;             // warps 0,1: (R,G); warps 2,3: (B,Y) - or swapped
    const int hidx = pair * S_OW + col;
#if JDS_SSIM_ST64
    // the two 8-byte halves of a 16-byte slot swap places in odd segments so that the 16 lanes
    // of a half-warp (two segments x eight rows) store to 16 different bank pairs
    const uint32_t h_swap = ((col >> 3) & 1) * 8;
#endif
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
    const bool p1_active = p1_row < S_R;
    const bool own_seg = S_SEG * p1_seg < own_px;            // own_px is a multiple of 16: whole segments
    double sse_a = 0.0, sse_b = 0.0;                // squared error of the pair's channels

    // prep: (B, Y) of both images, bytes -> centred fp32; task = 4 pixels of one (image, row).
    // Tasks tl, tl + ts, ... below t1.
#if JDS_SSIM_PREP_HOIST
    // the default split gives thread t the tasks t and t + 128 of every chunk: their source and
    // destination offsets never change, so the index arithmetic (two divisions by constants per
    // task) is done once here instead of in every chunk
    uint32_t prep_pk[2];                 // source byte offset | destination byte offset << 16
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int task = min(tid + k * S_NT, 2 * S_R * S_PG - 1);
        const int img = task / (S_R * S_PG);
        const int rem = task - img * (S_R * S_PG);
        const int r = rem / S_PG, g4 = rem - r * S_PG;
        const uint32_t dst = (uint32_t)(((img * S_R + r) * S_BPITCH + g4) * sizeof(float4));
        prep_pk[k] = (uint32_t)(img * S_TILEB + r * S_ROWB + 12 * g4) | (dst << 16);
    }
    const int prep_n = (tid + S_NT < 2 * S_R * S_PG) ? 2 : 1;
    auto prep_own = [&](int chunk) {
        const uint8_t* base = &sm.raw[chunk & 1][0][0];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k < prep_n) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(base + (prep_pk[k] & 0xffffu));
                float4* dst = reinterpret_cast<float4*>(reinterpret_cast<char*>(&sm.by[0][0][0][0]) + (prep_pk[k] >> 16));
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                const float R0 = byte_centered<0>(w0), G0 = byte_centered<1>(w0), B0 = byte_centered<2>(w0);
                const float R1 = byte_centered<3>(w0), G1 = byte_centered<0>(w1), B1 = byte_centered<1>(w1);
                const float R2 = byte_centered<2>(w1), G2 = byte_centered<3>(w1), B2 = byte_centered<0>(w2);
                const float R3 = byte_centered<1>(w2), G3 = byte_centered<2>(w2), B3 = byte_centered<3>(w2);
                const float Y0 = fmaf(0.299f, R0, fmaf(0.587f, G0, 0.114f * B0));
                const float Y1 = fmaf(0.299f, R1, fmaf(0.587f, G1, 0.114f * B1));
                const float Y2 = fmaf(0.299f, R2, fmaf(0.587f, G2, 0.114f * B2));
                const float Y3 = fmaf(0.299f, R3, fmaf(0.587f, G3, 0.114f * B3));
                dst[0] = make_float4(B0, Y0, B1, Y1);
                dst[S_BHALF] = make_float4(B2, Y2, B3, Y3);
            }
        }
    };
#endif
    auto prep_tasks = [&](int chunk, int tl, int t1, int ts) {
        const int buf = chunk & 1;
        for (int task = tl; task < t1; task += ts) {
            const int img = task / (S_R * S_PG);
            const int rem = task - img * (S_R * S_PG);
            const int r = rem / S_PG, g4 = rem - r * S_PG;
            const uint32_t* w = reinterpret_cast<const uint32_t*>(&sm.raw[buf][img][r * S_ROWB + 12 * g4]);
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
            sm.by[buf & (S_NBY - 1)][img][r][g4] = make_float4(B0, Y0, B1, Y1);
            sm.by[buf & (S_NBY - 1)][img][r][S_BHALF + g4] = make_float4(B2, Y2, B3, Y3);
        }
    };
    constexpr int N_PREP = 2 * S_R * S_PG;          // 252 tasks per chunk
#if JDS_SSIM_PREP_SPLIT
    // (B,Y) threads: tasks [0,128) after their pass 1 (two each) and [192,252) after pass 2;
    // (R,G) threads: tasks [128,192) after pass 2 (one each)
    constexpr int PREP_A = 128, PREP_B0 = 192;
#endif

    // chunk 0: everyone waits for the first tiles and prepares (B,Y)
    mbar_wait(&sm.bar[0], 0u);
#if JDS_SSIM_PREP_HOIST
    prep_own(0);
#else
    prep_tasks(0, tid, N_PREP, S_NT);
#endif
    __syncthreads();

    // one vertical step of pass 2: take row r's horizontal sums, emit (optionally), drop the
    // row that leaves the 7-row window
#if JDS_SSIM_ST64
    const uint32_t p2_xy = smem_u32(&sm.hxy[0][hidx]), p2_qc = smem_u32(&sm.hqc[0][hidx]);
#define JDS_P2_LOAD(r, h)                                                                \
    {                                                                                    \
        const uint32_t ro = (uint32_t)(r) * (S_HPITCH * 16);                             \
        h.sx = ld_shared_f2(p2_xy + ro + h_swap); h.sy = ld_shared_f2(p2_xy + ro + (8 - h_swap)); \
        h.sq = ld_shared_f2(p2_qc + ro + h_swap); h.sc = ld_shared_f2(p2_qc + ro + (8 - h_swap)); \
    }
#else
#define JDS_P2_LOAD(r, h)                                                                \
    {                                                                                    \
        const float4 h0 = sm.hxy[r][hidx], h1 = sm.hqc[r][hidx];                         \
        h.sx = f2(h0.x, h0.y); h.sy = f2(h0.z, h0.w);                                    \
        h.sq = f2(h1.x, h1.y); h.sc = f2(h1.z, h1.w);                                    \
    }
#endif
#define JDS_P2_STEP(r, EMIT)                                                             \
    {                                                                                    \
        HSum h;                                                                          \
        JDS_P2_LOAD(r, h)                                                                \
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
        const int buf = c & 1;
        const int y0 = py0 + c * S_R;
        const int nr = min(S_R, in_end - y0);

        // ---- pass 1: horizontal 7-tap sums of x, y, x^2+y^2, xy (sliding window) ---------
        if (p1_active) {
            const float2 sse = pair == 0 ? pass1_task<0>(sm, buf, p1_row, p1_seg)
                                         : pass1_task<1>(sm, buf, p1_row, p1_seg);
            if ((y0 + p1_row) < py1 && own_seg) {
                sse_a += (double)sse.x;
                sse_b += (double)sse.y;
            }
        }
#if JDS_SSIM_PREP_SPLIT
        if (pair == 1 && c + 1 < n_chunks) {
            // first part of the next chunk's (B,Y) planes (the other buffer): evens out the byte
            // conversion the (R,G) warps did above
            mbar_wait(&sm.bar[buf ^ 1], (uint32_t)(((c + 1) >> 1) & 1));
            prep_tasks(c + 1, tid - S_OW, PREP_A, S_OW);
        }
#endif
#if JDS_SSIM_PAIRBAR && !JDS_SSIM_PREP_SPLIT
        // hxy / hqc are private to a channel pair: only its own 64 threads (two warps) meet here,
        // so the (B,Y) warps never wait for the byte conversion of the (R,G) warps
        if (pair == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
        else asm volatile("bar.sync 2, 64;" ::: "memory");
#else
        __syncthreads();
#endif
        // raw[buf] (chunk c) has been consumed - by prep(c) before the last full barrier and by
        // the (R,G) threads just now (thread 0 is one of them): refill it with chunk c+2
        if (tid == 0 && c + 2 < n_chunks) issue(c + 2);

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
        if (c + 1 < n_chunks) {
#if JDS_SSIM_PREP_SPLIT
            if (pair == 0) {
                mbar_wait(&sm.bar[buf ^ 1], (uint32_t)(((c + 1) >> 1) & 1));
                prep_tasks(c + 1, PREP_A + tid, PREP_B0, S_OW);
            } else {
                prep_tasks(c + 1, PREP_B0 + tid - S_OW, N_PREP, S_OW);
            }
#elif JDS_SSIM_PAIRBAR
            // the (B,Y) warps take the larger share: the (R,G) warps converted bytes in pass 1
            mbar_wait(&sm.bar[buf ^ 1], (uint32_t)(((c + 1) >> 1) & 1));
            if (pair == 1) prep_tasks(c + 1, tid - S_OW, JDS_SSIM_PREP_BY1, S_OW);
            else prep_tasks(c + 1, JDS_SSIM_PREP_BY1 + tid, N_PREP, S_OW);
#else
            mbar_wait(&sm.bar[buf ^ 1], (uint32_t)(((c + 1) >> 1) & 1));
#if JDS_SSIM_PREP_RG > 0
            // the (R,G) warps converted bytes inside pass 1: the (B,Y) warps take more of the prep
            if (pair == 0) prep_tasks(c + 1, tid & (S_OW - 1), JDS_SSIM_PREP_RG, S_OW);
            else prep_tasks(c + 1, JDS_SSIM_PREP_RG + (tid & (S_OW - 1)), N_PREP, S_OW);
#elif JDS_SSIM_PREP_HOIST
            prep_own(c + 1);
#else
            prep_tasks(c + 1, tid, N_PREP, S_NT);
#endif
#endif
        }
        __syncthreads();
    }
#undef JDS_P2_STEP
#undef JDS_P2_LOAD

    // ---- reductions ---------------------------------------------------------------
    const int lane = tid & 31, warp = tid >> 5;
    // the pass-1 pair is warp uniform (warps 0,1: pair 0; warps 2,3: pair 1)
    double e[4];
    e[0] = (p1_active && pair == 0) ? sse_a : 0.0;
    e[1] = (p1_active && pair == 0) ? sse_b : 0.0;
    e[2] = (p1_active && pair == 1) ? sse_a : 0.0;
    e[3] = (p1_active && pair == 1) ? sse_b : 0.0;
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
        const int pr = tid >> 1, hf = tid & 1, w0 = 2 * (pr ^ swap);
        atomicAdd(&metrics[unit].ssim_sum[tid], 2.0 * (sm.red_ssim[w0][hf] + sm.red_ssim[w0 + 1][hf]));
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

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    // resolved per call: the lookup is a table probe inside the runtime and keeps this file free
    // of mutable statics (contexts on several threads / devices share nothing)
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    return reinterpret_cast<EncodeTiledFn>(fn);
}

// bytes x rows x units view of `units` packed uint8 RGB frames; unit stride 0 = one shared frame
static cudaError_t make_map(EncodeTiledFn enc, CUtensorMap* map, const uint8_t* base, int H, int W,
                            size_t unit_stride, int units) {
    const cuuint64_t dims[3] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)(unit_stride ? units : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 3, (cuuint64_t)(unit_stride ? unit_stride : (size_t)H * W * 3)};
    const cuuint32_t box[3] = {(cuuint32_t)S_ROWB, (cuuint32_t)S_R, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides,
                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// once per context (jds_ctx_create): dynamic shared memory opt-in and the largest carve-out so
// that four CTAs fit an SM; idempotent, no state shared between contexts
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
    const EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap map_a, map_b;
    cudaError_t e = make_map(enc, &map_a, a, H, W, a_stride, units);
    if (e == cudaSuccess) e = make_map(enc, &map_b, b, H, W, b_stride, units);
    if (e != cudaSuccess) return e;
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
        const double slots = (double)JDS_SSIM_MIN_CTAS * sm_count;
        const double work = (double)strips * units * H;
        double r_opt = sqrt(2.0 * 28.0 * work / slots);
        if (r_opt < 126.0) r_opt = 126.0;
        segs = (int)((double)H / r_opt);
        if (segs < 1) segs = 1;
        seg_rows = rows_for(segs);
        if (seg_rows < 126) seg_rows = 126;
    }
#endif
    if (const char* ov = getenv("JDS_SSIM_SEGS")) {      // A/B runs: force the vertical segment count
        const int n = atoi(ov);
        if (n >= 1) {
            seg_rows = rows_for(n);
            if (seg_rows < S_R) seg_rows = S_R;
        }
    }
    segs = (H + seg_rows - 1) / seg_rows;
    dim3 grid(strips, segs, units);
    k_ssim_strip<<<grid, S_NT, smem, s>>>(map_a, map_b, H, W, seg_rows, a_stride ? 1 : 0, b_stride ? 1 : 0,
                                          metrics, want_ssim ? 1 : 0, want_sse ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace jds
