// tc_boxsum.cu - does the tensor-core path beat the CUDA-core sliding window for SSIM's box sums?
// ("settle the tensor-core question with a measurement, not prose".)
//
// SSIM needs, per channel, the 7-tap window sums of x, y, x^2 + y^2 and xy (then the same again
// vertically).  Two stand-alone kernels produce the HORIZONTAL sums of one channel for the same
// centred int8 frames (x, y in [-128, 127], exactly what jds_ssim.cu works on):
//
//   k_tensor   band-matrix products on the tensor cores, integer exact.  Out = Data x Band with
//              mma.sync.m16n8k32 (s8 x s8 -> s32): the data rows are the A operand straight from
//              memory (row-major bytes ARE the A fragment layout), the 0/1 band matrix is a
//              constant B fragment; one 32-column K window yields 24 window sums (3 MMAs).
//              The products do not fit a byte, so q = x^2 + y^2 and c = xy go through the tensor
//              core as hi/lo byte planes: 6 byte planes per channel (x, y, q_hi, q_lo, c_hi,
//              c_lo).  The planes are PRE-SPLIT by an untimed kernel, i.e. the split (3 products
//              + 4 byte extractions per pixel on the CUDA cores) is not charged to this path.
//   k_sliding  the CUDA-core form of jds_ssim.cu's pass 1: a thread owns a row segment of 8
//              windows, converts 14 pixels, and slides the four sums (FADD / FFMA), products
//              included.
//
// Both write their sums (so the stores are comparable) and are timed with CUDA events on a frame
// set that stays in L2.  Reported: ms, window sums per clock per SM, and for the tensor kernel the
// MMA rate it achieved.  The SASS of k_tensor must show IMMA (cuobjdump -sass).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_boxsum tc_boxsum.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int H = 2160, W = 3840;          // one 4K channel
constexpr int NWIN = W - 6;                // window sums per row

// ---- pre-split (untimed): x, y -> the six byte planes of the tensor path -----------------------
__global__ void k_split(const int8_t* __restrict__ x, const int8_t* __restrict__ y, int8_t* __restrict__ planes,
                        size_t plane) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane) return;
    const int a = x[i], b = y[i];
    const int q = a * a + b * b - 16384, c = a * b;   // q (centred) in [-16384, 16384], c in [-16256, 16384]
    const int qh = (q + 128) >> 8, ql = q - 256 * qh; // q = 256 qh + ql, qh in [-64,64], ql in [-128,127]
    const int ch = (c + 128) >> 8, cl = c - 256 * ch; // c = 256 ch + cl
    planes[0 * plane + i] = (int8_t)a;
    planes[1 * plane + i] = (int8_t)b;
    planes[2 * plane + i] = (int8_t)qh;
    planes[3 * plane + i] = (int8_t)ql;
    planes[4 * plane + i] = (int8_t)ch;
    planes[5 * plane + i] = (int8_t)cl;
}

// ---- tensor-core band-matrix box sums ------------------------------------------------------------
__device__ __forceinline__ void imma_16832(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// a warp owns 16 rows of one byte plane and walks along them in steps of 24 window columns
__global__ void __launch_bounds__(128) k_tensor(const int8_t* __restrict__ planes, int n_planes, size_t plane,
                                                int* __restrict__ out) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int tiles_per_plane = H / 16;
    const int p = warp / tiles_per_plane, tile = warp % tiles_per_plane;
    if (p >= n_planes) return;
    // band fragments: B_j[k][n] = 1 for 0 <= k - (8 j + n) <= 6; this lane holds k = 4t..4t+3 (b0)
    // and k = 16+4t..16+4t+3 (b1) of column n = g
    uint32_t band[3][2];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t w = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = 16 * h + 4 * t + e, d = k - (8 * j + g);
                if (d >= 0 && d <= 6) w |= 1u << (8 * e);
            }
            band[j][h] = w;
        }
    const int8_t* src = planes + p * plane + (size_t)(tile * 16) * W;
    int* dst = out + (p * (size_t)H + tile * 16) * NWIN;
    for (int k0 = 0; k0 + 32 <= W; k0 += 24) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(src + (size_t)g * W + k0 + 4 * t);
        a[1] = *reinterpret_cast<const uint32_t*>(src + (size_t)(g + 8) * W + k0 + 4 * t);
        a[2] = *reinterpret_cast<const uint32_t*>(src + (size_t)g * W + k0 + 16 + 4 * t);
        a[3] = *reinterpret_cast<const uint32_t*>(src + (size_t)(g + 8) * W + k0 + 16 + 4 * t);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            int c[4] = {0, 0, 0, 0};
            imma_16832(c, a, band[j]);
            const int col = k0 + 8 * j + 2 * t;
            if (col + 1 < NWIN) {
                *reinterpret_cast<int2*>(dst + (size_t)g * NWIN + col) = make_int2(c[0], c[1]);
                *reinterpret_cast<int2*>(dst + (size_t)(g + 8) * NWIN + col) = make_int2(c[2], c[3]);
            }
        }
    }
}

// ---- CUDA-core sliding window (pass 1 of jds_ssim.cu for one channel, scalar fp32) ---------------
__global__ void __launch_bounds__(128) k_sliding(const int8_t* __restrict__ x, const int8_t* __restrict__ y,
                                                 float4* __restrict__ out) {
    // thread = (row, segment of 8 windows); 14 pixels per image: bytes 8 seg .. 8 seg + 13
    const int segs = NWIN / 8;                       // 479 (the last 2 window columns are left out on both sides)
    const size_t task = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (task >= (size_t)H * segs) return;
    const int row = (int)(task / segs), seg = (int)(task % segs);
    const uint2* px = reinterpret_cast<const uint2*>(x + (size_t)row * W + 8 * seg);
    const uint2* py = reinterpret_cast<const uint2*>(y + (size_t)row * W + 8 * seg);
    const uint2 xa = px[0], xb = px[1], ya = py[0], yb = py[1];
    const uint32_t wx[4] = {xa.x, xa.y, xb.x, xb.y}, wy[4] = {ya.x, ya.y, yb.x, yb.y};
    float xs[14], ys[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        // signed byte -> float: PRMT with sign replication into 2^23-biased form is not available
        // for signed data; the production kernel centres unsigned bytes, here the data is centred
        // already: one I2F-free conversion via the 0x4B000080 trick (value + 128 is unsigned)
        const uint32_t bx = (wx[i >> 2] >> (8 * (i & 3))) & 0xFFu, by = (wy[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        xs[i] = __uint_as_float(0x4B000000u | (bx ^ 0x80u)) - 8388736.0f;
        ys[i] = __uint_as_float(0x4B000000u | (by ^ 0x80u)) - 8388736.0f;
    }
    float sx = 0.f, sy = 0.f, sq = 0.f, sc = 0.f;
    float4* dst = out + (size_t)row * NWIN + 8 * seg;
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        sx += xs[i];
        sy += ys[i];
        sq = fmaf(xs[i], xs[i], fmaf(ys[i], ys[i], sq));
        sc = fmaf(xs[i], ys[i], sc);
        if (i >= 6) {
            const int j = i - 6;
            dst[j] = make_float4(sx, sy, sq, sc);
            if (j < 7) {
                sx -= xs[j];
                sy -= ys[j];
                sq = fmaf(-xs[j], xs[j], fmaf(-ys[j], ys[j], sq));
                sc = fmaf(-xs[j], ys[j], sc);
            }
        }
    }
}

// ---- raw IMMA issue rate (register operands only): the ceiling of the legacy tensor path ---------
__global__ void __launch_bounds__(256) k_imma_rate(int* out, int n) {
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u}, b[2] = {0x01010101u, 0x00010001u};
    int c[4][4] = {};
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) imma_16832(c[k], a, b);
    }
    int s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
    if (n < 0) out[threadIdx.x] = s;
}

int main() {
    const size_t plane = (size_t)H * W;
    int8_t *x, *y, *planes;
    int* out_t;
    float4* out_s;
    CK(cudaMalloc(&x, plane));
    CK(cudaMalloc(&y, plane));
    CK(cudaMalloc(&planes, 6 * plane));
    CK(cudaMalloc(&out_t, 6 * (size_t)H * NWIN * 4));
    CK(cudaMalloc(&out_s, (size_t)H * NWIN * 16));
    {
        int8_t* h = (int8_t*)malloc(2 * plane);
        srand(1);
        for (size_t i = 0; i < 2 * plane; ++i) h[i] = (int8_t)(rand() & 255);
        CK(cudaMemcpy(x, h, plane, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(y, h + plane, plane, cudaMemcpyHostToDevice));
        free(h);
    }
    k_split<<<(unsigned)((plane + 255) / 256), 256>>>(x, y, planes, plane);
    CK(cudaDeviceSynchronize());

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int reps = 20;
    float ms_t = 0, ms_s = 0;
    const int warps_t = 6 * (H / 16);
    for (int it = 0; it < 2; ++it) {                                   // first round warms up
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) k_tensor<<<(warps_t + 3) / 4, 128>>>(planes, 6, plane, out_t);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms_t, e0, e1);
        const size_t tasks = (size_t)H * (NWIN / 8);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) k_sliding<<<(unsigned)((tasks + 127) / 128), 128>>>(x, y, out_s);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms_s, e0, e1);
    }
    ms_t /= reps;
    ms_s /= reps;
    // check: the tensor sums against the sliding sums on one row
    {
        const int row = 777;
        int* ht = (int*)malloc(6 * (size_t)NWIN * 4);
        float4* hs = (float4*)malloc((size_t)NWIN * 16);
        for (int p = 0; p < 6; ++p)
            CK(cudaMemcpy(ht + (size_t)p * NWIN, out_t + ((size_t)p * H + row) * NWIN, (size_t)NWIN * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hs, out_s + (size_t)row * NWIN, (size_t)NWIN * 16, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int c = 0; c < (NWIN / 8) * 8 && c + 1 < 24 * ((W - 32) / 24 + 1); ++c) {
            const int sx = ht[c], sy = ht[NWIN + c];
            const int sq = 256 * ht[2 * NWIN + c] + ht[3 * NWIN + c] + 16384 * 7;
            const int sc = 256 * ht[4 * NWIN + c] + ht[5 * NWIN + c];
            if (sx != (int)hs[c].x || sy != (int)hs[c].y || sq != (int)hs[c].z || sc != (int)hs[c].w) ++bad;
        }
        printf("check row %d: %d mismatching window columns (tensor hi/lo recombined vs sliding fp32)\n", row, bad);
        free(ht);
        free(hs);
    }
    const double clk = 1.965e9, sms = 148.0;
    {
        const int n = 4096, grid = 148 * 8;
        k_imma_rate<<<grid, 256>>>(out_t, 16);
        cudaEventRecord(e0);
        k_imma_rate<<<grid, 256>>>(out_t, n);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double immas = (double)grid * 8 * n * 4;
        printf("raw IMMA.16832.S8 rate (registers only, 64 warps/SM): %.1f MACs/clk/SM = %.3f IMMA/clk/SM\n",
               immas * 4096 / (ms * 1e-3 * clk) / sms, immas / (ms * 1e-3 * clk) / sms);
    }
    const double chan_windows = (double)H * NWIN;                      // window sums of one channel (x4 quantities)
    const double mmas = (double)warps_t * ((W - 32) / 24 + 1) * 3;
    printf("tensor  (6 byte planes, IMMA m16n8k32): %.4f ms  -> %.2f channel-windows/clk/SM, %.0f MACs/clk/SM issued\n",
           ms_t, chan_windows / (ms_t * 1e-3 * clk) / sms, mmas * 16 * 8 * 32 / (ms_t * 1e-3 * clk) / sms);
    printf("sliding (fp32 CUDA cores, products incl.): %.4f ms  -> %.2f channel-windows/clk/SM\n", ms_s,
           chan_windows / (ms_s * 1e-3 * clk) / sms);
    printf("ratio tensor / sliding time: %.2f (the tensor path was NOT charged for splitting the products into bytes)\n",
           ms_t / ms_s);
    return 0;
}
