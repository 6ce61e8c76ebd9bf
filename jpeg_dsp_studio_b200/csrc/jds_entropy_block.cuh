// jds_entropy_block.cuh - the per-block part of the baseline-JPEG entropy coder (jds_entropy.cu),
// host + device so that tests/emul can run the very same code on the CPU against the oracle:
// the Annex K tables, canonical code construction, the zig-zag walk of one 8x8 block and the bit
// sinks it writes into.
#pragma once
#include <stdint.h>
#include "jds_math.cuh"

namespace jds {

// ITU-T T.81 Annex K.3 "typical" tables: BITS (codes per length 1..16) and HUFFVAL;
// rows of kBits: DC luminance, DC chrominance, AC luminance, AC chrominance
static const uint8_t kBits[4][16] = {
    {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},          // DC luminance
    {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0},          // DC chrominance
    {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d},       // AC luminance
    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77}};      // AC chrominance
static const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t kAcLumaVals[162] = {
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
static const uint8_t kAcChromaVals[162] = {
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


// canonical codes of T.81 Annex C, packed (code << 5) | length; 0 = the symbol has no code
inline void fill_packed(const uint8_t bits[16], const uint8_t* vals, int n_vals, uint32_t* out, int n_out) {
    for (int i = 0; i < n_out; ++i) out[i] = 0;
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int j = 0; j < bits[len - 1] && k < n_vals; ++j) out[vals[k++]] = (code++ << 5) | (uint32_t)len;
        code <<= 1;
    }
}


// (code << 5) | length of every symbol: dc[table][category], ac[table][run << 4 | size]
struct HuffPacked {
    uint32_t dc[2][12];
    uint32_t ac[2][256];
};
inline HuffPacked make_packed() {
    HuffPacked p;
    fill_packed(kBits[0], kDcVals, 12, p.dc[0], 12);
    fill_packed(kBits[1], kDcVals, 12, p.dc[1], 12);
    fill_packed(kBits[2], kAcLumaVals, 162, p.ac[0], 256);
    fill_packed(kBits[3], kAcChromaVals, 162, p.ac[1], 256);
    return p;
}

// utils/constants.py:18-27 ZIGZAG_ORDER, flattened: raster index of the k-th zig-zag coefficient
#define JDS_ZIGZAG_TABLE                                                                          \
    {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,     \
     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,     \
     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63}

JDS_HD int bit_size(int v) {      // SSSS: bits of |v|
    v = v < 0 ? -v : v;
#if defined(__CUDA_ARCH__)
    return 32 - __clz(v);
#else
    return v ? 32 - __builtin_clz((unsigned)v) : 0;
#endif
}

struct CountSink {
    unsigned int bits = 0;
    JDS_HD void put(uint32_t, int len) { bits += len; }
};

// appends bit strings (<= 27 bits each) at a bit position of a zero-initialised word array, first
// bit = most significant bit of word 0; words are shared with the neighbouring blocks, hence the
// OR (device: atomicOr on shared memory; host emulation: plain |=)
struct BitSink {
    uint32_t* words;
    unsigned long long acc = 0;
    int n;
    unsigned int wi;
    JDS_HD BitSink(uint32_t* w, unsigned int pos) : words(w), n((int)(pos & 31)), wi(pos >> 5) {}
    JDS_HD void or_word(unsigned int i, uint32_t v) {
#if defined(__CUDA_ARCH__)
        atomicOr(&words[i], v);
#else
        words[i] |= v;
#endif
    }
    JDS_HD void put(uint32_t code, int len) {
        acc = (acc << len) | code;
        n += len;
        if (n >= 32) {
            n -= 32;
            or_word(wi++, (uint32_t)(acc >> n));
            acc &= (1ull << n) - 1ull;
        }
    }
    JDS_HD void finish() {
        if (n) or_word(wi, (uint32_t)(acc << (32 - n)));
    }
};

// one block: DC difference, then the AC run/size symbols in zig-zag order.  dc / ac: packed
// (code << 5 | length) tables of the block's component; zz: the zig-zag table.  False when a value
// has no baseline code (DC difference beyond 11 bits, AC beyond 10).
template <class Sink>
JDS_HD bool walk_block(const int16_t* mine, int pred, const uint32_t* dc, const uint32_t* ac,
                       const uint8_t* zz, Sink& sink) {
    bool ok = true;
    const int d = (int)mine[0] - pred;
    const int dsz = bit_size(d);
    if (dsz > 11) {
        ok = false;
    } else {
        const uint32_t e = dc[dsz];
        const uint32_t amp = (uint32_t)(d >= 0 ? d : d + (1 << dsz) - 1);
        sink.put(((e >> 5) << dsz) | amp, (int)(e & 31) + dsz);
    }
    int run = 0;
    const uint32_t zrl = ac[0xF0];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = 1; k < 64; ++k) {
        const int v = mine[zz[k]];
        if (v == 0) {
            ++run;
            continue;
        }
        const int sz = bit_size(v);
        if (sz > 10) {
            ok = false;
            run = 0;
            continue;
        }
        while (run >= 16) {
            sink.put(zrl >> 5, (int)(zrl & 31));
            run -= 16;
        }
        const uint32_t e = ac[(run << 4) | sz];
        const uint32_t amp = (uint32_t)(v >= 0 ? v : v + (1 << sz) - 1);
        sink.put(((e >> 5) << sz) | amp, (int)(e & 31) + sz);
        run = 0;
    }
    if (run) sink.put(ac[0] >> 5, (int)(ac[0] & 31));
    return ok;
}

}  // namespace jds
