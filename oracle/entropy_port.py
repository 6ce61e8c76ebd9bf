"""Baseline-JPEG entropy coding of the round trip's coefficients (TEST INFRASTRUCTURE).

SURVEY.md 8f #4, second half: the reference only *estimates* a bit count
(``/root/reference/utils/metrics.py:51-92``, "no entropy coding") and defines
``ZIGZAG_ORDER`` (``utils/constants.py:18-27``) without using it.  This module states what a
real baseline JPEG (ITU-T T.81) spends on the same ``all_quantized_coeffs``
(``engines/pipeline.py:56,99``: channel Y|Cb|Cr -> block raster -> 64 row-major values):

  * zig-zag scan with the reference's ``ZIGZAG_ORDER`` table;
  * DC: difference to the previous block of the same component (raster order), category
    code + category bits; AC: (run, size) codes with ZRL and EOB;
  * the Annex K "typical" Huffman tables - luminance tables for Y, chrominance tables for
    Cb and Cr;
  * three non-interleaved scans (one per component), which is the block order the reference
    already stores, so the bit count needs no reordering.

``encode_jfif`` writes the complete file so that the arithmetic can be pinned to an
INDEPENDENT decoder: ``tests/test_entropy_cpu.py`` decodes it with OpenCV (libjpeg-turbo) and
compares with the round trip's own reconstruction, and compares the tables below with the ones
libjpeg-turbo itself writes.  There is no reference implementation to compare with - the
oracle of record here is "libjpeg decodes these bytes to the expected image".
"""

import numpy as np

#: /root/reference/utils/constants.py:18-27 - raster index of the k-th zig-zag coefficient
ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5,
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63], dtype=np.int64)

# ITU-T T.81 Annex K.3 "typical" tables: BITS (codes per length 1..16) and HUFFVAL
DC_LUMA_BITS = [0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
DC_LUMA_VALS = list(range(12))
DC_CHROMA_BITS = [0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
DC_CHROMA_VALS = list(range(12))
AC_LUMA_BITS = [0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d]
AC_LUMA_VALS = [
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
    0xf9, 0xfa]
AC_CHROMA_BITS = [0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77]
AC_CHROMA_VALS = [
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
    0xf9, 0xfa]

TABLES = {("dc", 0): (DC_LUMA_BITS, DC_LUMA_VALS), ("dc", 1): (DC_CHROMA_BITS, DC_CHROMA_VALS),
          ("ac", 0): (AC_LUMA_BITS, AC_LUMA_VALS), ("ac", 1): (AC_CHROMA_BITS, AC_CHROMA_VALS)}


def build_codes(bits, vals):
    """T.81 Annex C: canonical codes.  Returns (code[256], length[256]); length 0 = no code."""
    code = np.zeros(256, dtype=np.int64)
    length = np.zeros(256, dtype=np.int64)
    c, k = 0, 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            code[vals[k]], length[vals[k]] = c, ln
            c += 1
            k += 1
        c <<= 1
    assert k == len(vals)
    return code, length


CODES = {k: build_codes(*v) for k, v in TABLES.items()}


def plane_blocks(shape, mode):
    """Blocks (rows, cols) of Y and of each chroma plane, as the pipeline lays them out
    (engines/color_space.py:44-49 sizes, engines/block_processor.py:7-16 padding)."""
    h, w = shape
    hc, wc = {"4:4:4": (h, w), "4:2:2": (h, w // 2), "4:2:0": (h // 2, w // 2)}[mode]
    nb = lambda n: (n + 7) // 8
    return (nb(h), nb(w)), (nb(hc), nb(wc)), (hc, wc)


def split_components(coeffs, shape, mode):
    """all_quantized_coeffs -> [Y, Cb, Cr] arrays of shape (n_blocks, 64), block raster order."""
    (by, bx), (cy, cx), _ = plane_blocks(shape, mode)
    ny, nc = by * bx, cy * cx
    c = np.asarray(coeffs).astype(np.int64).ravel()
    assert c.size == 64 * (ny + 2 * nc), (c.size, ny, nc)
    return [c[:64 * ny].reshape(ny, 64), c[64 * ny:64 * (ny + nc)].reshape(nc, 64),
            c[64 * (ny + nc):].reshape(nc, 64)]


def _category(v):
    """SSSS: number of bits of |v| (0 for 0)."""
    a = np.abs(v)
    out = np.zeros(a.shape, dtype=np.int64)
    nz = a > 0
    out[nz] = np.floor(np.log2(a[nz])).astype(np.int64) + 1
    return out


def component_scan_bits(blocks, table_id):
    """Exact entropy-coded bits of one non-interleaved scan (before byte stuffing / padding)."""
    _, dc_len = CODES[("dc", table_id)]
    _, ac_len = CODES[("ac", table_id)]
    zz = blocks[:, ZIGZAG]
    dc = zz[:, 0]
    diff = np.diff(dc, prepend=0)
    cat = _category(diff)
    bits = int((dc_len[cat] + cat).sum())
    ac = zz[:, 1:]
    for b in range(ac.shape[0]):
        row = ac[b]
        nzpos = np.nonzero(row)[0]
        prev = -1
        for p in nzpos:
            run = p - prev - 1
            bits += (run // 16) * int(ac_len[0xF0])
            size = int(_category(row[p:p + 1])[0])
            bits += int(ac_len[((run % 16) << 4) | size]) + size
            prev = p
        if prev != 62:
            bits += int(ac_len[0x00])
    return bits


def huffman_scan_bits(coeffs, shape, mode):
    """[bits_Y, bits_Cb, bits_Cr] - what the three scans of a baseline JPEG spend."""
    comps = split_components(coeffs, shape, mode)
    return [component_scan_bits(comps[0], 0), component_scan_bits(comps[1], 1),
            component_scan_bits(comps[2], 1)]


# ---------------------------------------------------------------------------
# complete JFIF file (validation against an independent decoder)
# ---------------------------------------------------------------------------
class _BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0
        self.payload_bits = 0

    def put(self, code, length):
        self.payload_bits += length
        self.acc = (self.acc << length) | (code & ((1 << length) - 1))
        self.n += length
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0x00)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            pad = 8 - self.n
            self.acc = (self.acc << pad) | ((1 << pad) - 1)
            self.n = 8
            b = self.acc & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0x00)
            self.acc = self.n = 0


def _encode_scan(blocks, table_id):
    dc_code, dc_len = CODES[("dc", table_id)]
    ac_code, ac_len = CODES[("ac", table_id)]
    w = _BitWriter()
    pred = 0

    def amplitude(v, size):
        return v if v >= 0 else v + (1 << size) - 1

    for blk in blocks:
        zz = blk[ZIGZAG]
        d = int(zz[0]) - pred
        pred = int(zz[0])
        s = int(abs(d)).bit_length()
        w.put(int(dc_code[s]), int(dc_len[s]))
        if s:
            w.put(amplitude(d, s), s)
        run = 0
        for k in range(1, 64):
            v = int(zz[k])
            if v == 0:
                run += 1
                continue
            while run >= 16:
                w.put(int(ac_code[0xF0]), int(ac_len[0xF0]))
                run -= 16
            s = abs(v).bit_length()
            sym = (run << 4) | s
            w.put(int(ac_code[sym]), int(ac_len[sym]))
            w.put(amplitude(v, s), s)
            run = 0
        if run:
            w.put(int(ac_code[0x00]), int(ac_len[0x00]))
    bits = w.payload_bits
    w.flush()
    return bytes(w.out), bits


def encode_jfif(coeffs, shape, mode, qtable):
    """Baseline sequential JPEG, three non-interleaved scans.  ``qtable``: the 8x8 table the
    round trip used (the reference uses the luminance table for all three components,
    engines/pipeline.py:43).  Needs even W under 4:2:2 / 4:2:0 and even H under 4:2:0 so that
    JPEG's ceil() component sizes equal the pipeline's floor().  Returns (bytes, scan_bits)."""
    h, w = shape
    if mode != "4:4:4" and (w % 2 or (mode == "4:2:0" and h % 2)):
        raise ValueError("odd frame sizes: JPEG's component geometry differs from the pipeline's")
    comps = split_components(coeffs, shape, mode)
    q = np.asarray(qtable).astype(np.int64).ravel()
    assert q.min() >= 1 and q.max() <= 255
    hs, vs = {"4:4:4": (1, 1), "4:2:2": (2, 1), "4:2:0": (2, 2)}[mode]

    def seg(marker, payload):
        return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload

    out = bytearray(b"\xFF\xD8")
    out += seg(0xE0, b"JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00")
    out += seg(0xDB, bytes([0x00]) + bytes(int(q[i]) for i in ZIGZAG))
    sof = bytes([8]) + h.to_bytes(2, "big") + w.to_bytes(2, "big") + bytes([3])
    sof += bytes([1, (hs << 4) | vs, 0, 2, 0x11, 0, 3, 0x11, 0])
    out += seg(0xC0, sof)
    for (kind, tid), (bits, vals) in TABLES.items():
        out += seg(0xC4, bytes([(0x10 if kind == "ac" else 0x00) | tid]) + bytes(bits) + bytes(vals))
    scan_bits = []
    for ci, blocks in enumerate(comps):
        tid = 0 if ci == 0 else 1
        out += seg(0xDA, bytes([1, ci + 1, (tid << 4) | tid, 0, 63, 0]))
        data, nbits = _encode_scan(blocks, tid)
        out += data
        scan_bits.append(nbits)
    out += b"\xFF\xD9"
    return bytes(out), scan_bits


# ---------------------------------------------------------------------------
# decoder (round-trip check of the coder above; libjpeg is the independent check)
# ---------------------------------------------------------------------------
def decode_scan(data, n_blocks, table_id):
    """Inverse of ``_encode_scan``: entropy-coded bytes of one scan -> (n_blocks, 64) raster
    coefficients.  Removes the 0xFF00 stuffing, walks the canonical codes bit by bit."""
    raw = bytearray()
    i = 0
    while i < len(data):
        raw.append(data[i])
        if data[i] == 0xFF:
            assert data[i + 1] == 0x00, "marker inside entropy-coded data"
            i += 1
        i += 1
    bits = np.unpackbits(np.frombuffer(bytes(raw), dtype=np.uint8))
    pos = 0

    def lookup(kind):
        code, length = CODES[(kind, table_id)]
        table = {(int(length[s]), int(code[s])): s for s in range(256) if length[s]}
        return table

    dc_t, ac_t = lookup("dc"), lookup("ac")

    def symbol(table):
        nonlocal pos
        c = 0
        for ln in range(1, 17):
            c = (c << 1) | int(bits[pos])
            pos += 1
            if (ln, c) in table:
                return table[(ln, c)]
        raise ValueError("invalid Huffman code")

    def amplitude(size):
        nonlocal pos
        if size == 0:
            return 0
        v = 0
        for _ in range(size):
            v = (v << 1) | int(bits[pos])
            pos += 1
        return v if v >= (1 << (size - 1)) else v - (1 << size) + 1

    out = np.zeros((n_blocks, 64), dtype=np.int64)
    pred = 0
    for b in range(n_blocks):
        pred += amplitude(symbol(dc_t))
        out[b, ZIGZAG[0]] = pred
        k = 1
        while k < 64:
            s = symbol(ac_t)
            run, size = s >> 4, s & 15
            if size == 0:
                if run == 15:
                    k += 16
                    continue
                break                          # EOB
            k += run
            out[b, ZIGZAG[k]] = amplitude(size)
            k += 1
    return out, pos
