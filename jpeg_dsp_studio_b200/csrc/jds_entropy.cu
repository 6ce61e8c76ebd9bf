// jds_entropy.cu - exact size of the baseline-JPEG entropy-coded scans of the round trip's
// coefficients (second half).  The reference only estimates a bit count
// ("no entropy coding", utils/metrics.py:51-92) and defines ZIGZAG_ORDER
// (utils/constants.py:18-27) without using it; this is what ITU-T T.81 Huffman coding with
// the Annex K "typical" tables spends on the same int16 array (engines/pipeline.py:56,99:
// channel Y|Cb|Cr -> block raster -> 64 row-major values), coded as three non-interleaved
// scans - the block order the array already has:
//   DC  code(category of diff to the previous block of the component) + category bits
//   AC  zig-zag order; per non-zero: ZRL for every 16 zeros of the run, code(run%16, size)
//       + size bits; EOB when the block ends in zeros
// One thread owns one 8x8 block: its 128 bytes are staged in shared memory (vector loads,
// conflict-free padded slots) and walked in zig-zag order; code lengths come from shared
// memory copies of the four tables.  HBM-bound: 128 B read per block, three atomics per CTA.
#include <cuda_runtime.h>
#include <stdint.h>
#include "jds_kernels.cuh"

namespace jds {

namespace {

// code LENGTHS of the Annex K tables (T.81 K.3-K.6), index = symbol; generated from BITS/HUFFVAL
struct HuffLengths {
    uint8_t dc[2][12];
    uint8_t ac[2][256];
};

const uint8_t kBits[4][16] = {
    {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},          // DC luminance
    {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0},          // DC chrominance
    {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d},       // AC luminance
    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77}};      // AC chrominance
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07,
    0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0,
    0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49,
    0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7,
    0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5,
    0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa};
const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71,
    0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0,
    0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5,
    0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa};

void fill_lengths(const uint8_t bits[16], const uint8_t* vals, int n_vals, uint8_t* out, int n_out) {
    for (int i = 0; i < n_out; ++i) out[i] = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len)
        for (int j = 0; j < bits[len - 1] && k < n_vals; ++j) out[vals[k++]] = (uint8_t)len;
}

HuffLengths make_lengths() {
    HuffLengths h;
    fill_lengths(kBits[0], kDcVals, 12, h.dc[0], 12);
    fill_lengths(kBits[1], kDcVals, 12, h.dc[1], 12);
    fill_lengths(kBits[2], kAcLumaVals, 162, h.ac[0], 256);
    fill_lengths(kBits[3], kAcChromaVals, 162, h.ac[1], 256);
    return h;
}

}  // namespace

__constant__ HuffLengths c_huff;
// utils/constants.py:18-27 ZIGZAG_ORDER, flattened: raster index of the k-th zig-zag coefficient
__constant__ uint8_t c_zigzag[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int E_NT = 128;
constexpr int E_SLOT = 66;          // int16 per staged block: 33 words (odd) -> lane i sits in bank i + const

__device__ __forceinline__ int bit_size(int v) {      // SSSS: bits of |v|
    v = v < 0 ? -v : v;
    return 32 - __clz(v);
}

// blocks [0, ny) are Y, [ny, ny+nc) Cb, [ny+nc, ny+2nc) Cr; scan_bits[3]
__global__ void __launch_bounds__(E_NT)
k_entropy_bits(const int16_t* __restrict__ coeffs, long long ny, long long nc,
               unsigned long long* __restrict__ scan_bits) {
    __shared__ __align__(4) int16_t stage[E_NT * E_SLOT];
    __shared__ uint8_t s_ac[2][256];
    __shared__ uint8_t s_dc[2][12];
    __shared__ unsigned long long s_sum[3];
    const int tid = threadIdx.x;
    for (int i = tid; i < 512; i += E_NT) s_ac[i >> 8][i & 255] = c_huff.ac[i >> 8][i & 255];
    if (tid < 24) s_dc[tid / 12][tid % 12] = c_huff.dc[tid / 12][tid % 12];
    if (tid < 3) s_sum[tid] = 0ull;
    const long long total = ny + 2 * nc;
    const long long b = (long long)blockIdx.x * E_NT + tid;
    const bool live = b < total;
    int comp = 0, pred = 0;
    int16_t* mine = stage + tid * E_SLOT;
    if (live) {
        comp = b < ny ? 0 : (b < ny + nc ? 1 : 2);
        const long long first = comp == 0 ? 0 : (comp == 1 ? ny : ny + nc);
        const int4* src = reinterpret_cast<const int4*>(coeffs + b * 64);
        // 8 x 16-byte loads; the slot is 4-byte aligned (odd word pitch), so store words
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int4 v = __ldg(src + i);
            int* dst = reinterpret_cast<int*>(mine + 8 * i);
            dst[0] = v.x;
            dst[1] = v.y;
            dst[2] = v.z;
            dst[3] = v.w;
        }
        if (b > first) pred = (int)__ldg(coeffs + (b - 1) * 64);
    }
    __syncthreads();
    unsigned int bits = 0;
    if (live) {
        const int tb = comp ? 1 : 0;
        const int dsz = bit_size((int)mine[0] - pred);
        bits = s_dc[tb][dsz] + dsz;
        int run = 0;
#pragma unroll 1
        for (int k = 1; k < 64; ++k) {
            const int v = mine[c_zigzag[k]];
            if (v == 0) {
                ++run;
            } else {
                const int sz = bit_size(v);
                bits += (run >> 4) * s_ac[tb][0xF0] + s_ac[tb][((run & 15) << 4) | sz] + sz;
                run = 0;
            }
        }
        if (run) bits += s_ac[tb][0];
    }
    // per-component sums: a warp may straddle a component boundary, so reduce by atomics on
    // the CTA's three shared counters, then one global atomic per component
    if (live) atomicAdd(&s_sum[comp], (unsigned long long)bits);
    __syncthreads();
    if (tid < 3 && s_sum[tid]) atomicAdd(&scan_bits[tid], s_sum[tid]);
}

// once per context (jds_ctx_create): the Annex K code lengths in constant memory of this device
cudaError_t entropy_configure_device() {
    const HuffLengths h = make_lengths();
    return cudaMemcpyToSymbol(c_huff, &h, sizeof h);
}

cudaError_t launch_entropy_bits(const int16_t* coeffs, long long ny, long long nc,
                                unsigned long long* scan_bits, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(scan_bits, 0, 3 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    const long long total = ny + 2 * nc;
    k_entropy_bits<<<(unsigned)((total + E_NT - 1) / E_NT), E_NT, 0, s>>>(coeffs, ny, nc, scan_bits);
    return cudaGetLastError();
}

}  // namespace jds
