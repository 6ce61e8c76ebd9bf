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
//
// The BITSTREAM of the same scans is produced on the device too (jds_entropy_encode /
// jds_jfif_encode): per-block sizes -> exclusive prefix sums -> every block writes its codes at
// its own bit offset (a CTA assembles its 64 blocks in shared memory and stores whole words,
// coalesced) -> 0xFF bytes get their stuffed 0x00 by a second count / scan / scatter pass.
#include <cuda_runtime.h>
#include <stdint.h>
#include "jds_kernels.cuh"
#include "jds_entropy_block.cuh"

namespace jds {

namespace {

// code LENGTHS of the Annex K tables (T.81 K.3-K.6), index = symbol; generated from BITS/HUFFVAL
struct HuffLengths {
    uint8_t dc[2][12];
    uint8_t ac[2][256];
};

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
__constant__ uint8_t c_zigzag[64] = JDS_ZIGZAG_TABLE;

constexpr int E_NT = 128;
constexpr int E_SLOT = 66;          // int16 per staged block: 33 words (odd) -> lane i sits in bank i + const


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

// ---------------------------------------------------------------------------------------
// The bitstream.  Every scan (component) owns a run of CTAs, 64 blocks each (EntropyGrid), so a
// CTA's output is one contiguous bit range of one scan.
//   1. k_entropy_block_bits   bits of every block + the sum per CTA
//   2. k_entropy_layout       exclusive prefix of the CTA sums per scan; scan sizes; where each
//                             scan starts in the unstuffed buffer (4-byte aligned)
//   3. k_entropy_pack         codes of the CTA's blocks OR-ed into a shared-memory image of its
//                             bit range (big-endian bit order, T.81 F.1.2.3), stored as whole
//                             words; only the two boundary words need a global atomic.  The last
//                             block of a scan appends the 1-bits that pad the final byte.
//   4. k_stuff_count / k_stuff_layout / k_stuff_scatter   T.81 B.1.1.5: 0x00 after every 0xFF
//      byte - a stream compaction with the bytes between the scans' 4-byte slots dropped, so the
//      three scans come out back to back.
// ---------------------------------------------------------------------------------------
__constant__ HuffPacked c_pack;

constexpr int P_NT = ENTROPY_CTA_BLOCKS;
constexpr int P_MAX_BLOCK_BITS = 22 + 63 * 26;   // DC 11 + 11; 63 x (16-bit code + 10 bits)
constexpr int P_WORDS = (31 + P_NT * P_MAX_BLOCK_BITS + 7 + 31) / 32 + 1;

struct BlockSlot {
    int comp;
    long long local, global;
    bool live;
};

__device__ __forceinline__ BlockSlot locate(const EntropyGrid& g, int cta, int tid) {
    BlockSlot b;
    b.comp = cta >= g.cta0[2] ? 2 : (cta >= g.cta0[1] ? 1 : 0);
    b.local = (long long)(cta - g.cta0[b.comp]) * P_NT + tid;
    b.live = b.local < g.n[b.comp];
    b.global = g.first[b.comp] + b.local;
    return b;
}

// the block's 128 bytes into its padded shared-memory slot, the tables of its component, and the
// DC predictor (previous block of the scan)
__device__ __forceinline__ int stage_block(const int16_t* __restrict__ coeffs, const BlockSlot& b,
                                           int16_t* mine, uint32_t* s_dc, uint32_t* s_ac, int tid) {
    const int tb = b.comp ? 1 : 0;
    for (int i = tid; i < 256; i += P_NT) s_ac[i] = c_pack.ac[tb][i];
    if (tid < 12) s_dc[tid] = c_pack.dc[tb][tid];
    int pred = 0;
    if (b.live) {
        const int4* src = reinterpret_cast<const int4*>(coeffs + b.global * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int4 v = __ldg(src + i);
            int* dst = reinterpret_cast<int*>(mine + 8 * i);
            dst[0] = v.x;
            dst[1] = v.y;
            dst[2] = v.z;
            dst[3] = v.w;
        }
        if (b.local > 0) pred = (int)__ldg(coeffs + (b.global - 1) * 64);
    }
    return pred;
}

// exclusive prefix of one value per thread over the CTA (NT threads, NT/32 <= 32 warps);
// *total = the CTA's sum.  Contains barriers: call from all threads.
template <int NT, class T>
__device__ __forceinline__ T cta_exclusive_scan(T v, T* s_warp, T* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const T y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        T w = lane < NT / 32 ? s_warp[lane] : (T)0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const T y = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += y;
        }
        if (lane < NT / 32) s_warp[lane] = w;
    }
    __syncthreads();
    const T before = warp ? s_warp[warp - 1] : (T)0;
    *total = s_warp[NT / 32 - 1];
    __syncthreads();                       // s_warp may be reused by the caller's next scan
    return before + x - v;
}

__global__ void __launch_bounds__(P_NT)
k_entropy_block_bits(const int16_t* __restrict__ coeffs, EntropyGrid g, uint32_t* __restrict__ blk_bits,
                     uint32_t* __restrict__ part, EntropyLayout* __restrict__ lay) {
    __shared__ __align__(4) int16_t stage[P_NT * E_SLOT];
    __shared__ uint32_t s_ac[256], s_dc[12], s_warp[P_NT / 32];
    const int tid = threadIdx.x;
    const BlockSlot b = locate(g, blockIdx.x, tid);
    int16_t* mine = stage + tid * E_SLOT;
    const int pred = stage_block(coeffs, b, mine, s_dc, s_ac, tid);
    __syncthreads();
    CountSink sink;
    if (b.live) {
        if (!walk_block(mine, pred, s_dc, s_ac, c_zigzag, sink)) atomicOr(&lay->invalid, 1u);
        blk_bits[b.global] = sink.bits;
    }
    uint32_t total;
    cta_exclusive_scan<P_NT, uint32_t>(sink.bits, s_warp, &total);
    if (tid == 0) part[blockIdx.x] = total;
}

// exclusive prefix of in[lo, hi) into out[lo, hi) by ONE CTA of 1024 threads; returns the sum
// (valid in every thread)
__device__ unsigned long long scan_range_1024(const uint32_t* __restrict__ in, long long lo, long long hi,
                                              unsigned long long* __restrict__ out,
                                              unsigned long long* s_warp) {
    unsigned long long carry = 0;
    for (long long base = lo; base < hi; base += 1024) {
        const long long i = base + threadIdx.x;
        const unsigned long long v = i < hi ? in[i] : 0ull;
        unsigned long long total;
        const unsigned long long ex = cta_exclusive_scan<1024, unsigned long long>(v, s_warp, &total);
        if (i < hi) out[i] = carry + ex;
        carry += total;
    }
    return carry;
}

__global__ void __launch_bounds__(1024)
k_entropy_layout(const uint32_t* __restrict__ part, EntropyGrid g, unsigned long long* __restrict__ cta_off,
                 EntropyLayout* __restrict__ lay) {
    __shared__ unsigned long long s_warp[32];
    unsigned long long bits[3];
    for (int comp = 0; comp < 3; ++comp)
        bits[comp] = scan_range_1024(part, g.cta0[comp], comp < 2 ? g.cta0[comp + 1] : g.ctas, cta_off, s_warp);
    if (threadIdx.x == 0) {
        unsigned long long at = 0;
        for (int comp = 0; comp < 3; ++comp) {
            lay->bits[comp] = bits[comp];
            lay->ubytes[comp] = (bits[comp] + 7) >> 3;
            lay->ubase[comp] = at;
            at = (at + lay->ubytes[comp] + 3) & ~3ull;
        }
        lay->total_ubytes = at;
    }
}

__global__ void __launch_bounds__(P_NT)
k_entropy_pack(const int16_t* __restrict__ coeffs, EntropyGrid g, const uint32_t* __restrict__ blk_bits,
               const unsigned long long* __restrict__ cta_off, const EntropyLayout* __restrict__ lay,
               uint32_t* __restrict__ ubuf) {
    __shared__ __align__(4) int16_t stage[P_NT * E_SLOT];
    __shared__ uint32_t s_ac[256], s_dc[12], s_warp[P_NT / 32];
    __shared__ uint32_t words[P_WORDS];
    const int tid = threadIdx.x;
    const BlockSlot b = locate(g, blockIdx.x, tid);
    int16_t* mine = stage + tid * E_SLOT;
    const int pred = stage_block(coeffs, b, mine, s_dc, s_ac, tid);
    // the last block of the scan also writes the 1-bits that complete the final byte
    const bool last = b.live && b.local == g.n[b.comp] - 1;
    const int pad = last ? (int)((8 - (lay->bits[b.comp] & 7)) & 7) : 0;
    const uint32_t mybits = b.live ? blk_bits[b.global] + pad : 0u;
    uint32_t cta_bits;
    const uint32_t before = cta_exclusive_scan<P_NT, uint32_t>(mybits, s_warp, &cta_bits);   // syncs
    const unsigned long long start = lay->ubase[b.comp] * 8ull + cta_off[blockIdx.x];
    const unsigned int shift = (unsigned int)(start & 31);
    const unsigned int n_words = (shift + cta_bits + 31) >> 5;
    for (unsigned int i = tid; i < n_words; i += P_NT) words[i] = 0u;
    __syncthreads();
    if (b.live) {
        BitSink sink(words, shift + before);
        walk_block(mine, pred, s_dc, s_ac, c_zigzag, sink);
        if (pad) sink.put((1u << pad) - 1u, pad);
        sink.finish();
    }
    __syncthreads();
    uint32_t* dst = ubuf + (start >> 5);
    for (unsigned int i = tid; i < n_words; i += P_NT) {
        const uint32_t v = __byte_perm(words[i], 0, 0x0123);       // first bit = MSB of byte 0
        if (i == 0 || i == n_words - 1) {
            if (v) atomicOr(dst + i, v);
        } else {
            dst[i] = v;
        }
    }
}

// ---- byte stuffing ----------------------------------------------------------------------
constexpr int S_NT = 256;
constexpr int S_CHUNK = S_NT * 16;

// of the 16 bytes at i0: bit j set when byte i0+j belongs to scan `comp`
__device__ __forceinline__ unsigned int scan_mask(const EntropyLayout& L, int comp, unsigned long long i0) {
    const unsigned long long a = L.ubase[comp], e = a + L.ubytes[comp];
    const unsigned long long lo = i0 > a ? i0 : a, hi = i0 + 16 < e ? i0 + 16 : e;
    if (hi <= lo) return 0u;
    return ((1u << (unsigned int)(hi - lo)) - 1u) << (unsigned int)(lo - i0);
}

__device__ __forceinline__ unsigned int ff_mask(const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    unsigned int m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (((w[k] >> (8 * j)) & 0xFFu) == 0xFFu) m |= 1u << (4 * k + j);
    return m;
}

__global__ void __launch_bounds__(S_NT)
k_stuff_count(const uint4* __restrict__ ubuf, EntropyLayout* __restrict__ lay, uint32_t* __restrict__ cnt) {
    __shared__ uint32_t s_warp[S_NT / 32];
    __shared__ unsigned int s_ff[3];
    __shared__ EntropyLayout L;
    if (threadIdx.x == 0) L = *lay;
    if (threadIdx.x < 3) s_ff[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned long long idx = (unsigned long long)blockIdx.x * S_NT + threadIdx.x;
    const unsigned long long i0 = idx * 16;
    uint32_t emit = 0;
    if (i0 < L.total_ubytes) {
        const unsigned int ff = ff_mask(__ldg(ubuf + idx));
        for (int comp = 0; comp < 3; ++comp) {
            const unsigned int m = scan_mask(L, comp, i0);
            const unsigned int f = __popc(m & ff);
            emit += __popc(m) + f;
            if (f) atomicAdd(&s_ff[comp], f);
        }
    }
    uint32_t total;
    cta_exclusive_scan<S_NT, uint32_t>(emit, s_warp, &total);
    if (threadIdx.x == 0) cnt[blockIdx.x] = total;
    if (threadIdx.x < 3 && s_ff[threadIdx.x]) atomicAdd(&lay->ff[threadIdx.x], (unsigned long long)s_ff[threadIdx.x]);
}

__global__ void __launch_bounds__(1024)
k_stuff_layout(const uint32_t* __restrict__ cnt, long long n_chunks, unsigned long long* __restrict__ chunk_off,
               EntropyLayout* __restrict__ lay) {
    __shared__ unsigned long long s_warp[32];
    const unsigned long long total = scan_range_1024(cnt, 0, n_chunks, chunk_off, s_warp);
    if (threadIdx.x == 0) lay->stuffed_bytes = total;
}

__global__ void __launch_bounds__(S_NT)
k_stuff_scatter(const uint4* __restrict__ ubuf, const EntropyLayout* __restrict__ lay,
                const unsigned long long* __restrict__ chunk_off, uint8_t* __restrict__ out) {
    __shared__ uint32_t s_warp[S_NT / 32];
    __shared__ EntropyLayout L;
    __shared__ uint8_t s_out[2 * S_CHUNK];
    if (threadIdx.x == 0) L = *lay;
    __syncthreads();
    const unsigned long long idx = (unsigned long long)blockIdx.x * S_NT + threadIdx.x;
    const unsigned long long i0 = idx * 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    unsigned int m = 0, ff = 0;
    if (i0 < L.total_ubytes) {
        v = __ldg(ubuf + idx);
        ff = ff_mask(v);
        m = scan_mask(L, 0, i0) | scan_mask(L, 1, i0) | scan_mask(L, 2, i0);
        ff &= m;
    }
    uint32_t total;
    uint32_t at = cta_exclusive_scan<S_NT, uint32_t>(__popc(m) + __popc(ff), s_warp, &total);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (m & (1u << j)) {
            s_out[at++] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
            if (ff & (1u << j)) s_out[at++] = 0;
        }
    }
    __syncthreads();
    uint8_t* dst = out + chunk_off[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < total; i += S_NT) dst[i] = s_out[i];
}

// once per context (jds_ctx_create): the Annex K code lengths in constant memory of this device
cudaError_t entropy_configure_device() {
    const HuffLengths h = make_lengths();
    cudaError_t e = cudaMemcpyToSymbol(c_huff, &h, sizeof h);
    if (e != cudaSuccess) return e;
    const HuffPacked p = make_packed();
    return cudaMemcpyToSymbol(c_pack, &p, sizeof p);
}

cudaError_t launch_entropy_bits(const int16_t* coeffs, long long ny, long long nc,
                                unsigned long long* scan_bits, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(scan_bits, 0, 3 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    const long long total = ny + 2 * nc;
    k_entropy_bits<<<(unsigned)((total + E_NT - 1) / E_NT), E_NT, 0, s>>>(coeffs, ny, nc, scan_bits);
    return cudaGetLastError();
}

EntropyGrid make_entropy_grid(long long ny, long long nc) {
    EntropyGrid g;
    const long long n[3] = {ny, nc, nc};
    long long first = 0;
    int cta = 0;
    for (int k = 0; k < 3; ++k) {
        g.n[k] = n[k];
        g.first[k] = first;
        g.cta0[k] = cta;
        first += n[k];
        cta += (int)((n[k] + P_NT - 1) / P_NT);
    }
    g.ctas = cta;
    return g;
}

// steps 1-2: sizes and layout (device).  blk_bits: one word per block; part / cta_off: one
// entry per CTA of the grid; lay: zeroed here.
cudaError_t launch_entropy_sizes(const int16_t* coeffs, const EntropyGrid& g, uint32_t* blk_bits,
                                 uint32_t* part, unsigned long long* cta_off, EntropyLayout* lay,
                                 cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(lay, 0, sizeof(EntropyLayout), s);
    if (e != cudaSuccess) return e;
    k_entropy_block_bits<<<g.ctas, P_NT, 0, s>>>(coeffs, g, blk_bits, part, lay);
    k_entropy_layout<<<1, 1024, 0, s>>>(part, g, cta_off, lay);
    return cudaGetLastError();
}

// step 3: ubuf = ubuf_bytes (a multiple of 4096, >= lay->total_ubytes) zeroed here
cudaError_t launch_entropy_pack(const int16_t* coeffs, const EntropyGrid& g, const uint32_t* blk_bits,
                                const unsigned long long* cta_off, const EntropyLayout* lay,
                                uint32_t* ubuf, size_t ubuf_bytes, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(ubuf, 0, ubuf_bytes, s);
    if (e != cudaSuccess) return e;
    k_entropy_pack<<<g.ctas, P_NT, 0, s>>>(coeffs, g, blk_bits, cta_off, lay, ubuf);
    return cudaGetLastError();
}

// step 4a: emitted bytes per 4096-byte chunk, 0xFF bytes per scan, stuffed size
cudaError_t launch_stuff_sizes(const uint32_t* ubuf, size_t ubuf_bytes, EntropyLayout* lay, uint32_t* cnt,
                               unsigned long long* chunk_off, cudaStream_t s) {
    const long long chunks = (long long)(ubuf_bytes / S_CHUNK);
    if (chunks) k_stuff_count<<<(unsigned)chunks, S_NT, 0, s>>>(reinterpret_cast<const uint4*>(ubuf), lay, cnt);
    k_stuff_layout<<<1, 1024, 0, s>>>(cnt, chunks, chunk_off, lay);
    return cudaGetLastError();
}

// step 4b: the stuffed scans, back to back, into out (lay->stuffed_bytes bytes)
cudaError_t launch_stuff_scatter(const uint32_t* ubuf, size_t ubuf_bytes, const EntropyLayout* lay,
                                 const unsigned long long* chunk_off, uint8_t* out, cudaStream_t s) {
    const long long chunks = (long long)(ubuf_bytes / S_CHUNK);
    if (chunks)
        k_stuff_scatter<<<(unsigned)chunks, S_NT, 0, s>>>(reinterpret_cast<const uint4*>(ubuf), lay, chunk_off, out);
    return cudaGetLastError();
}

// The fixed part of a baseline JFIF file for the round trip's coefficients: SOI, APP0, one DQT
// (the reference quantises all three components with the luminance table,
// engines/pipeline.py:43), SOF0, the four Annex K DHT segments.  Returns the byte count;
// out may be NULL (size only).  sub: 0 4:4:4, 1 4:2:2, 2 4:2:0.
size_t jfif_write_headers(uint8_t* out, int height, int width, int sub, const uint8_t q_raster[64]) {
    static const uint8_t zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                   12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                   58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    size_t n = 0;
    auto put = [&](int b) {
        if (out) out[n] = (uint8_t)b;
        ++n;
    };
    auto seg = [&](int marker, int payload) {
        put(0xFF);
        put(marker);
        put((payload + 2) >> 8);
        put((payload + 2) & 255);
    };
    put(0xFF);
    put(0xD8);
    seg(0xE0, 14);
    for (int b : {0x4A, 0x46, 0x49, 0x46, 0, 1, 1, 0, 0, 1, 0, 1, 0, 0}) put(b);   // "JFIF\0" 1.01, 1:1
    seg(0xDB, 65);
    put(0);
    for (int k = 0; k < 64; ++k) put(q_raster[zz[k]]);
    seg(0xC0, 15);
    put(8);
    put(height >> 8);
    put(height & 255);
    put(width >> 8);
    put(width & 255);
    put(3);
    const int hs = sub == 0 ? 1 : 2, vs = sub == 2 ? 2 : 1;
    for (int b : {1, (hs << 4) | vs, 0, 2, 0x11, 0, 3, 0x11, 0}) put(b);
    const uint8_t* vals[4] = {kDcVals, kDcVals, kAcLumaVals, kAcChromaVals};
    const int n_vals[4] = {12, 12, 162, 162};
    const int ident[4] = {0x00, 0x01, 0x10, 0x11};
    for (int t = 0; t < 4; ++t) {
        seg(0xC4, 17 + n_vals[t]);
        put(ident[t]);
        for (int k = 0; k < 16; ++k) put(kBits[t][k]);
        for (int k = 0; k < n_vals[t]; ++k) put(vals[t][k]);
    }
    return n;
}

// SOS header of the non-interleaved scan of component `comp` (0 Y, 1 Cb, 2 Cr): 10 bytes
size_t jfif_write_sos(uint8_t* out, int comp) {
    const int tid = comp ? 1 : 0;
    const uint8_t h[10] = {0xFF, 0xDA, 0, 8, 1, (uint8_t)(comp + 1), (uint8_t)((tid << 4) | tid), 0, 63, 0};
    if (out)
        for (int k = 0; k < 10; ++k) out[k] = h[k];
    return 10;
}

}  // namespace jds
