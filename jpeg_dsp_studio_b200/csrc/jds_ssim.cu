// jds_ssim.cu - strip-streaming SSIM / SSE kernel (sm_100a).
//
// Computes, for R, G, B and BT.601 Y of two uint8 RGB images (original,
// reconstructed): the sum of the 7x7-window SSIM map over all windows inside the
// image (utils/metrics.py:12-14,21 -> skimage structural_similarity), and the sums
// of squared differences for PSNR (utils/metrics.py:11,20).
//
// Design (DESIGN.md "k_ssim_strip"):
//   * a CTA owns a vertical strip of 128 window columns (134 pixel columns) and walks
//     down it 7 rows at a time; nothing is recomputed vertically.
//   * rows arrive by TMA bulk copies (cp.async.bulk + mbarrier), double buffered.
//   * pass 1 (horizontal): a thread owns (row, channel, 16-window segment) and forms
//     the 7-tap window sums of x, y, x^2+y^2, xy by running prefix differences in
//     registers; results go to shared memory as one float4 per (row, channel, column).
//   * pass 2 (vertical): a thread owns (column, channel), keeps the last 7 horizontal
//     sums in a register ring, slides the 7-row sum and evaluates the SSIM formula.
//   * samples are centred (x-128) so fp32 sums of the integer channels are exact and
//     the fp32 Y sums keep their low bits; the Y accumulators are rebuilt from the
//     ring every chunk so rounding cannot drift down a strip.
#include <cuda_runtime.h>
#include <stdint.h>
#include "jds_kernels.cuh"

namespace jds {

constexpr int S_OW = 128;            // window columns per strip
constexpr int S_LW = 144;            // pixel columns loaded per strip (multiple of 16)
constexpr int S_ROWB = S_LW * 3;     // bytes per loaded row
constexpr int S_R = 7;               // rows per chunk == window height
constexpr int S_NT = 512;            // threads per CTA
constexpr int S_SEG = 16;            // windows per pass-1 task
constexpr int S_NSEG = S_OW / S_SEG; // 8

struct SsimSmem {
    alignas(128) uint8_t raw[2][2][S_R][S_ROWB];   // [buffer][image][row][byte]
    alignas(16) float ybuf[2][S_R][S_LW];          // centred Y of both images
    alignas(16) float4 hbuf[S_R][4][S_OW];         // horizontal window sums (swizzled)
    alignas(8) unsigned long long bar[2];
    double red_ssim[4][16];
    double red_ssey[16];
    unsigned long long red_sse[16];
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

__device__ __forceinline__ int hswz(int col) { return col ^ ((col >> 4) & 7); }

// byte `b` (0..3) of `w` as (value - 128) in fp32: 0x4B0000vv is 2^23 + vv
template <int B>
__device__ __forceinline__ float byte_centered(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | B)) - 8388736.0f;
}

// SSIM of one window from centred sums over its 49 samples
__device__ __forceinline__ float ssim_window(float sx, float sy, float sq, float sxy) {
    constexpr float inv = 1.0f / 49.0f;
    constexpr float cn = 49.0f / 48.0f;
    constexpr float C1 = 6.5025f, C2 = 58.5225f;
    const float mx = sx * inv, my = sy * inv;
    const float t = mx * my;
    const float A2 = fmaf(sxy, 2.0f * cn * inv, fmaf(t, -2.0f * cn, C2));
    const float u = fmaf(mx, mx, my * my);
    const float B2 = fmaf(sq, cn * inv, fmaf(u, -cn, C2));
    const float ux = mx + 128.0f, uy = my + 128.0f;
    const float A1 = fmaf(ux * uy, 2.0f, C1);
    const float B1 = fmaf(ux, ux, fmaf(uy, uy, C1));
    return __fdividef(A1 * A2, B1 * B2);
}

// ---- pass 1 for one (row, segment) of channel CH ---------------------------------
template <int CH>
__device__ __forceinline__ void pass1_task(SsimSmem& sm, int buf, int r, int seg, int valid_px,
                                           float& sse_out) {
    float px[S_SEG + 6], py[S_SEG + 6];
    if (CH < 3) {
        const uint4* qa = reinterpret_cast<const uint4*>(&sm.raw[buf][0][r][48 * seg]);
        const uint4* qb = reinterpret_cast<const uint4*>(&sm.raw[buf][1][r][48 * seg]);
        uint32_t wa[20], wb[20];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            uint4 a = qa[i], b = qb[i];
            wa[4 * i] = a.x; wa[4 * i + 1] = a.y; wa[4 * i + 2] = a.z; wa[4 * i + 3] = a.w;
            wb[4 * i] = b.x; wb[4 * i + 1] = b.y; wb[4 * i + 2] = b.z; wb[4 * i + 3] = b.w;
        }
#pragma unroll
        for (int i = 0; i < S_SEG + 6; ++i) {
            constexpr int dummy = 0;
            (void)dummy;
            const int b = 3 * i + CH;
            const uint32_t va = wa[b >> 2], vb = wb[b >> 2];
            switch (b & 3) {
                case 0: px[i] = byte_centered<0>(va); py[i] = byte_centered<0>(vb); break;
                case 1: px[i] = byte_centered<1>(va); py[i] = byte_centered<1>(vb); break;
                case 2: px[i] = byte_centered<2>(va); py[i] = byte_centered<2>(vb); break;
                default: px[i] = byte_centered<3>(va); py[i] = byte_centered<3>(vb); break;
            }
        }
    } else {
        const float4* qa = reinterpret_cast<const float4*>(&sm.ybuf[0][r][S_SEG * seg]);
        const float4* qb = reinterpret_cast<const float4*>(&sm.ybuf[1][r][S_SEG * seg]);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            float4 a = qa[i], b = qb[i];
            if (4 * i + 0 < S_SEG + 6) { px[4 * i + 0] = a.x; py[4 * i + 0] = b.x; }
            if (4 * i + 1 < S_SEG + 6) { px[4 * i + 1] = a.y; py[4 * i + 1] = b.y; }
            if (4 * i + 2 < S_SEG + 6) { px[4 * i + 2] = a.z; py[4 * i + 2] = b.z; }
            if (4 * i + 3 < S_SEG + 6) { px[4 * i + 3] = a.w; py[4 * i + 3] = b.w; }
        }
    }
    // squared error of the pixels this task owns (its first 16 columns)
    float sse = 0.f;
    const int c0 = S_SEG * seg;
    if (c0 + S_SEG <= valid_px) {
#pragma unroll
        for (int i = 0; i < S_SEG; ++i) {
            const float d = px[i] - py[i];
            sse = fmaf(d, d, sse);
        }
    } else {
#pragma unroll
        for (int i = 0; i < S_SEG; ++i) {
            const float d = (c0 + i < valid_px) ? px[i] - py[i] : 0.f;
            sse = fmaf(d, d, sse);
        }
    }
    sse_out += sse;
    // running prefix sums; window j = prefix[j+7] - prefix[j]
    float Px[S_SEG + 7], Py[S_SEG + 7], Pq[S_SEG + 7], Pc[S_SEG + 7];
    Px[0] = Py[0] = Pq[0] = Pc[0] = 0.f;
#pragma unroll
    for (int i = 0; i < S_SEG + 6; ++i) {
        Px[i + 1] = Px[i] + px[i];
        Py[i + 1] = Py[i] + py[i];
        Pq[i + 1] = fmaf(px[i], px[i], fmaf(py[i], py[i], Pq[i]));
        Pc[i + 1] = fmaf(px[i], py[i], Pc[i]);
    }
#pragma unroll
    for (int j = 0; j < S_SEG; ++j) {
        float4 h;
        h.x = Px[j + 7] - Px[j];
        h.y = Py[j + 7] - Py[j];
        h.z = Pq[j + 7] - Pq[j];
        h.w = Pc[j + 7] - Pc[j];
        sm.hbuf[r][CH][hswz(c0 + j)] = h;
    }
}

__global__ void __launch_bounds__(S_NT, 1)
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
    // rows: this CTA owns pixel rows [py0, py1) for the squared error, and the window
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
        // one warp issues the chunk's bulk copies (14 rows), lane i -> row i
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
    if (tid < 32) issue(0);

    // pass-2 state: thread = (column, channel)
    const int col = tid & (S_OW - 1);
    const int ch = tid >> 7;
    float4 ring[S_R];
#pragma unroll
    for (int i = 0; i < S_R; ++i) ring[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    double ssim_total = 0.0;
    const bool col_ok = col < nwin_x;
    // pass-1 state: warps 0..7, two warps per channel, 56 tasks per channel
    const int p1_ch = tid >> 6;
    const int p1_task = tid & 63;
    const bool p1_active = tid < 256 && p1_task < S_R * S_NSEG;
    const int p1_row = p1_task / S_NSEG, p1_seg = p1_task % S_NSEG;
    double sse_total = 0.0;

    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        if (tid < 32 && c + 1 < n_chunks) issue(c + 1);
        mbar_wait(&sm.bar[buf], (uint32_t)((c >> 1) & 1));
        const int y0 = py0 + c * S_R;
        const int nr = min(S_R, in_end - y0);

        // ---- centred Y of both images: one task = 4 pixels -------------------------
        if (tid < 2 * S_R * (S_LW / 4)) {
            const int img = tid / (S_R * (S_LW / 4));
            const int rem = tid % (S_R * (S_LW / 4));
            const int r = rem / (S_LW / 4), g4 = rem % (S_LW / 4);
            if (r < nr) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(&sm.raw[buf][img][r][12 * g4]);
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                float4 yv;
                yv.x = fmaf(0.299f, byte_centered<0>(w0), fmaf(0.587f, byte_centered<1>(w0), 0.114f * byte_centered<2>(w0)));
                yv.y = fmaf(0.299f, byte_centered<3>(w0), fmaf(0.587f, byte_centered<0>(w1), 0.114f * byte_centered<1>(w1)));
                yv.z = fmaf(0.299f, byte_centered<2>(w1), fmaf(0.587f, byte_centered<3>(w1), 0.114f * byte_centered<0>(w2)));
                yv.w = fmaf(0.299f, byte_centered<1>(w2), fmaf(0.587f, byte_centered<2>(w2), 0.114f * byte_centered<3>(w2)));
                *reinterpret_cast<float4*>(&sm.ybuf[img][r][4 * g4]) = yv;
            }
        }
        __syncthreads();

        // ---- pass 1: horizontal window sums --------------------------------------
        if (p1_active && p1_row < nr) {
            // squared error only for rows this CTA owns (not the 6-row overlap)
            float sse = 0.f;
            const bool own_row = (y0 + p1_row) < py1;
            const int vpx = own_row ? own_px : 0;
            switch (p1_ch) {
                case 0: pass1_task<0>(sm, buf, p1_row, p1_seg, vpx, sse); break;
                case 1: pass1_task<1>(sm, buf, p1_row, p1_seg, vpx, sse); break;
                case 2: pass1_task<2>(sm, buf, p1_row, p1_seg, vpx, sse); break;
                default: pass1_task<3>(sm, buf, p1_row, p1_seg, vpx, sse); break;
            }
            sse_total += (double)sse;
        }
        __syncthreads();

        // ---- pass 2: vertical sliding sum + SSIM ------------------------------------
        if (want_ssim) {
            if (ch == 3) {
                // rebuild the fp32 accumulators from the ring: no drift down the strip
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 1; i < S_R; ++i) {
                    acc.x += ring[i].x; acc.y += ring[i].y; acc.z += ring[i].z; acc.w += ring[i].w;
                }
            }
            float ssum = 0.f;
#pragma unroll
            for (int r = 0; r < S_R; ++r) {
                if (r < nr) {
                    const float4 h = sm.hbuf[r][ch][hswz(col)];
                    ring[r] = h;
                    acc.x += h.x; acc.y += h.y; acc.z += h.z; acc.w += h.w;
                    const int step = c * S_R + r;            // rows consumed so far - 1
                    const int wy = py0 + step - 6;           // window row that just completed
                    if (step >= 6 && wy < py1 && col_ok)
                        ssum += ssim_window(acc.x, acc.y, acc.z, acc.w);
                    const float4 o = ring[(r + 1) % S_R];
                    acc.x -= o.x; acc.y -= o.y; acc.z -= o.z; acc.w -= o.w;
                }
            }
            ssim_total += (double)ssum;
        }
        __syncthreads();
    }

    // ---- reductions ---------------------------------------------------------------
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ssim_total += __shfl_down_sync(0xffffffffu, ssim_total, o);
        sse_total += __shfl_down_sync(0xffffffffu, sse_total, o);
    }
    if (lane == 0) {
        sm.red_ssim[0][warp] = ssim_total;     // warp -> channel = warp / 4
        sm.red_ssey[warp] = sse_total;         // pass-1 warps: channel = warp / 2 (warps 0..7)
    }
    __syncthreads();
    if (tid < 4) {
        double s = 0.0;
        for (int w = 0; w < 4; ++w) s += sm.red_ssim[0][tid * 4 + w];
        if (want_ssim) atomicAdd(&metrics[unit].ssim_sum[tid], s);
    }
    if (tid == 32 && want_sse) {
        // integer channels: every partial is an exact integer below 2^53
        double rgb = 0.0;
        for (int w = 0; w < 6; ++w) rgb += sm.red_ssey[w];
        atomicAdd(&metrics[unit].sse_rgb, (unsigned long long)(rgb + 0.5));
        atomicAdd(&metrics[unit].sse_y, sm.red_ssey[6] + sm.red_ssey[7]);
    }
}

size_t ssim_strip_smem_bytes() { return sizeof(SsimSmem); }

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
    {
        cudaError_t e = cudaFuncSetAttribute(k_ssim_strip, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int strips = (W + S_OW - 1) / S_OW;
    // vertical segments: enough CTAs to fill the machine (~2 per SM x 2 waves), but at
    // least 126 rows each so the 6-row overlap stays below 5 %
    int want_ctas = sm_count * 4;
    int segs = (want_ctas + strips * units - 1) / (strips * units);
    if (segs < 1) segs = 1;
    int seg_rows = (H + segs - 1) / segs;
    if (seg_rows < 126) seg_rows = 126;
    seg_rows = (seg_rows + S_R - 1) / S_R * S_R;
    segs = (H + seg_rows - 1) / seg_rows;
    dim3 grid(strips, segs, units);
    k_ssim_strip<<<grid, S_NT, smem, s>>>(H, W, seg_rows, a, a_stride, b, b_stride, metrics,
                                          want_ssim ? 1 : 0, want_sse ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace jds
