"""Quantisation - the reference's ``engines/quantizer.py:7-29`` API.  The table scaling is
the library's host routine (``jds_quant_table``, integer-exact against the reference
formula for every quality); quantise / dequantise run on the GPU in exact arithmetic
(IEEE divide, round-half-even, int16)."""

import ctypes as C

import numpy as np

from .. import _native as N
from ..engine import get_engine
from ..utils.constants import JPEG_LUMA_Q50

_QUANTIZE, _DEQUANTIZE = 4, 5


def scale_quant_matrix(base_matrix: np.ndarray, quality: int) -> np.ndarray:
    """IJG quality scaling of a quantisation table (quantizer.py:7-19)."""
    quality = int(np.clip(quality, 1, 100))
    if base_matrix is JPEG_LUMA_Q50 or np.array_equal(base_matrix, JPEG_LUMA_Q50):
        t = (C.c_double * 64)()
        N.check(N.load().jds_quant_table(quality, t))
        return np.array(t, dtype=np.float64).reshape(8, 8)
    # any other base table: the same formula, evaluated on the host (64 values)
    scale = 5000.0 / quality if quality < 50 else 200.0 - 2.0 * quality
    return np.clip(np.floor((np.asarray(base_matrix, dtype=np.float64) * scale + 50.0) / 100.0),
                   1, 255).astype(np.float64)


def _check(blocks, Q):
    a = np.asarray(blocks)
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    if a.shape[-2:] != Q.shape or Q.shape != (8, 8):
        # same failure as the reference's broadcast of (B,B) against (8,8)
        raise ValueError(f"operands could not be broadcast together with shapes "
                         f"{tuple(a.shape[-2:])} {tuple(Q.shape)} ")
    return Q


def quantize(dct_coeffs: np.ndarray, Q_matrix: np.ndarray) -> np.ndarray:
    """``np.round(dct / Q).astype(int16)`` (quantizer.py:22-24)."""
    Q = _check(dct_coeffs, Q_matrix)
    a = np.ascontiguousarray(dct_coeffs, dtype=np.float64)
    out = np.empty(a.shape, dtype=np.int16)
    eng = get_engine()
    with eng._lock:
        N.check(eng._lib.jds_block_op(eng._ctx, _QUANTIZE, a.size // 64, C.c_void_p(a.ctypes.data),
                                      None, C.c_void_p(Q.ctypes.data), None,
                                      C.c_void_p(out.ctypes.data)))
    return out


def dequantize(quantized: np.ndarray, Q_matrix: np.ndarray) -> np.ndarray:
    """``quantized.astype(float64) * Q`` (quantizer.py:27-29)."""
    Q = _check(quantized, Q_matrix)
    a = np.ascontiguousarray(quantized, dtype=np.int16)
    out = np.empty(a.shape, dtype=np.float64)
    eng = get_engine()
    with eng._lock:
        N.check(eng._lib.jds_block_op(eng._ctx, _DEQUANTIZE, a.size // 64, None,
                                      C.c_void_p(a.ctypes.data), C.c_void_p(Q.ctypes.data),
                                      C.c_void_p(out.ctypes.data), None))
    return out
