"""Entropy coding (second half) without a GPU: the oracle's baseline-JPEG coder
(oracle/entropy_port.py) is pinned to an INDEPENDENT implementation - libjpeg-turbo through
OpenCV: its Huffman tables equal the ones libjpeg writes, and the complete JFIF file it emits
decodes to the round trip's own reconstruction."""

import numpy as np
import pytest

from oracle import entropy_port as E
from oracle import numpy_port as P
from jpeg_dsp_studio_b200.utils import constants as K
from jpeg_dsp_studio_b200.utils import test_images as TI
from jpeg_dsp_studio_b200.utils import metrics as M


def test_zigzag_is_the_reference_table():
    assert np.array_equal(E.ZIGZAG, K.ZIGZAG_ORDER.ravel())
    assert sorted(E.ZIGZAG.tolist()) == list(range(64))


def test_tables_are_complete_prefix_codes():
    for (kind, tid), (bits, vals) in E.TABLES.items():
        assert sum(bits) == len(vals) == len(set(vals))
        code, length = E.CODES[(kind, tid)]
        words = sorted((int(length[v]), int(code[v])) for v in vals)
        strs = [format(c, "0%db" % n) for n, c in words]
        for i, a in enumerate(strs):                      # no code is a prefix of another
            assert not any(b.startswith(a) for b in strs[i + 1:])
        assert sum(2.0 ** -n for n, _ in words) < 1.0    # all-ones code word stays reserved


def test_tables_equal_libjpeg_turbo():
    cv2 = pytest.importorskip("cv2")
    ok, buf = cv2.imencode(".jpg", np.random.default_rng(0).integers(0, 256, (64, 64, 3), dtype=np.uint8))
    b, i, found = buf.tobytes(), 2, {}
    while i < len(b):
        marker, size = b[i + 1], int.from_bytes(b[i + 2:i + 4], "big")
        if marker == 0xC4:
            p = b[i + 4:i + 2 + size]
            while p:
                n = sum(p[1:17])
                found[("ac" if p[0] >> 4 else "dc", p[0] & 15)] = (list(p[1:17]), list(p[17:17 + n]))
                p = p[17 + n:]
        if marker == 0xDA:
            break
        i += 2 + size
    for k, (bits, vals) in E.TABLES.items():
        assert found[k] == (list(bits), list(vals)), k


@pytest.mark.parametrize("mode", ["4:4:4", "4:2:2", "4:2:0"])
@pytest.mark.parametrize("q", [5, 50, 95])
def test_jfif_decodes_to_the_round_trip(mode, q):
    """libjpeg's integer IDCT / upsampling / colour conversion differ from the reference's
    float pipeline by a few grey levels at most; a wrong code length, run or sign would give
    garbage (or no image at all)."""
    cv2 = pytest.importorskip("cv2")
    for img in (TI.generate_photo(256)[:96, :144], np.random.default_rng(1).integers(0, 256, (40, 56, 3), dtype=np.uint8)):
        o = P.compress_reconstruct(img, q, mode, False, want_maps=False)
        data, bits = E.encode_jfif(o["all_quantized_coeffs"], img.shape[:2], mode, P.scale_quant_matrix(q))
        assert bits == E.huffman_scan_bits(o["all_quantized_coeffs"], img.shape[:2], mode)
        dec = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert dec is not None and dec.shape == img.shape
        d = dec[..., ::-1].astype(np.int32) - o["reconstructed_image"].astype(np.int32)
        assert np.abs(d).max() <= 4
        assert 10 * np.log10(255.0 ** 2 / max(float(np.mean(d * d)), 1e-12)) >= 45.0
        # the file is the fixed headers + the padded scans + 0xFF stuffing bytes
        payload = sum((b + 7) // 8 for b in bits)
        stuffing = len(data) - M.JFIF_HEADER_BYTES - payload
        assert 0 <= stuffing <= 0.05 * payload + 4


def test_extreme_values_and_runs():
    """largest magnitudes (Q=100 table of ones: DC +-1016..1024, category 11), a lone last
    coefficient (three ZRL + run 14) and an empty block (EOB only)"""
    blocks = np.zeros((4, 64), dtype=np.int64)
    blocks[0, 0] = -1024
    blocks[1, 0] = 1016                       # diff 2040 -> category 11
    blocks[1, 63] = -1                        # zig-zag position 63: run 62 = 3 ZRL + 14
    blocks[2, E.ZIGZAG[1]] = 1023             # size 10
    _, dcl = E.CODES[("dc", 0)]
    _, acl = E.CODES[("ac", 0)]
    want = (dcl[11] + 11 + acl[0]) + (dcl[11] + 11 + 3 * acl[0xF0] + acl[(14 << 4) | 1] + 1) + \
           (dcl[10] + 10 + acl[0x0A] + 10 + acl[0]) + (dcl[0] + acl[0])
    assert E.component_scan_bits(blocks, 0) == int(want)
    data, nbits = E._encode_scan(blocks, 0)
    assert nbits == int(want)


def test_coder_round_trips_its_own_bitstream():
    """encode -> decode returns the coefficients exactly and consumes exactly the counted bits"""
    rng = np.random.default_rng(11)
    for table_id in (0, 1):
        blocks = np.zeros((40, 64), dtype=np.int64)
        dense = rng.integers(-1023, 1024, (40, 64))
        keep = rng.random((40, 64)) < rng.choice([0.02, 0.2, 0.9], size=(40, 1))
        blocks[keep] = dense[keep]
        blocks[:, 0] = rng.integers(-1024, 1017, 40)
        blocks[7] = 0
        blocks[8, 63] = 5                                   # zig-zag tail after a long run
        data, nbits = E._encode_scan(blocks, table_id)
        back, used = E.decode_scan(data, 40, table_id)
        assert np.array_equal(back, blocks)
        assert used == nbits == E.component_scan_bits(blocks, table_id)


# ---------------------------------------------------------------------------------------
# the product's per-block coder (csrc/jds_entropy_block.cuh: code tables, zig-zag walk, bit sink)
# compiled for the host (tests/emul) against the oracle coder above - the CUDA kernels run the
# same source per thread (tests/test_entropy_gpu.py checks them on the device)
# ---------------------------------------------------------------------------------------
def _emul_scan(lib, blocks, table_id):
    import ctypes as C
    blocks = np.ascontiguousarray(blocks, dtype=np.int16)
    cap = 4 * blocks.size + 64
    out = np.zeros(cap, dtype=np.uint8)
    nbytes, nbits = C.c_uint64(), C.c_uint64()
    rc = lib.emul_entropy_scan(blocks.ctypes.data_as(C.c_void_p), C.c_longlong(blocks.shape[0]), table_id,
                               out.ctypes.data_as(C.c_void_p), C.c_uint64(cap), C.byref(nbytes), C.byref(nbits))
    return rc, bytes(out[:nbytes.value]), int(nbits.value)


def test_device_block_coder_on_cpu_matches_oracle():
    import ctypes as C
    from tests.emul import build as EB
    lib = C.CDLL(EB.build())
    lib.emul_entropy_scan.restype = C.c_int
    rng = np.random.default_rng(17)
    cases = []
    for n, spread, keep in ((1, 30, 64), (7, 300, 20), (64, 40, 10), (65, 1023, 64), (200, 8, 3), (33, 0, 0)):
        b = rng.integers(-spread, spread + 1, (n, 64)).astype(np.int16)
        b[:, keep:] = 0
        cases.append(b)
    ext = np.zeros((12, 64), dtype=np.int16)
    ext[:, 0] = [-1024, 1016, 0, 5, -1024, 1016, 1016, -1024, 0, 0, 7, 7]
    ext[1, 63] = -1                                   # three ZRL then one coefficient
    ext[2, 1] = 1023
    ext[5] = 1023                                     # all-ones amplitudes: 0xFF bytes to stuff
    ext[6] = -1023
    ext[9, 62] = 77
    cases.append(ext)
    for blocks in cases:
        for tid in (0, 1):
            want, want_bits = E._encode_scan(blocks.astype(np.int64), tid)
            rc, got, bits = _emul_scan(lib, blocks, tid)
            assert rc == 0 and bits == want_bits and got == want, (blocks.shape, tid)
    bad = ext.copy()
    bad[3, 5] = 1024                                  # no baseline code: refused
    assert _emul_scan(lib, bad, 0)[0] == -1
    bad = ext.copy()
    bad[3, 0], bad[4, 0] = 2047, -2047                # DC difference of 12 bits
    assert _emul_scan(lib, bad, 1)[0] == -1
