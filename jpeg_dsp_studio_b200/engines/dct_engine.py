"""DCT / IDCT with level shift - the reference's ``engines/dct_engine.py:7-27`` API on the
GPU (exact mode: bit-identical to ``scipy.fft.dctn/idctn(type=2, norm='ortho')`` as the
reference calls them).  Accepts one 8x8 block or any ``(..., 8, 8)`` stack of blocks."""

import ctypes as C

import numpy as np

from .. import _native as N
from ..engine import get_engine

_DCT2, _IDCT2, _ENCODE, _DECODE = 0, 1, 2, 3


def _block_op(op: int, blocks: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(blocks, dtype=np.float64)
    if a.ndim < 2 or a.shape[-2:] != (8, 8):
        # the reference accepts any 2-D block for dct2/idct2; the round trip only ever uses 8x8
        raise ValueError(f"expected (..., 8, 8) blocks, got shape {a.shape}")
    out = np.empty_like(a)
    eng = get_engine()
    with eng._lock:
        N.check(eng._lib.jds_block_op(eng._ctx, op, a.size // 64, C.c_void_p(a.ctypes.data), None,
                                      None, C.c_void_p(out.ctypes.data), None))
    return out


def dct2(block: np.ndarray) -> np.ndarray:
    """2-D DCT-II with orthonormal normalisation (dct_engine.py:7-9)."""
    return _block_op(_DCT2, block)


def idct2(coeffs: np.ndarray) -> np.ndarray:
    """2-D inverse DCT, type III (dct_engine.py:12-14)."""
    return _block_op(_IDCT2, coeffs)


def encode_block(block: np.ndarray) -> np.ndarray:
    """Level shift (-128) then DCT (dct_engine.py:17-20)."""
    return _block_op(_ENCODE, block)


def decode_block(coeffs: np.ndarray) -> np.ndarray:
    """IDCT, +128, clip to [0, 255] (dct_engine.py:23-27)."""
    return _block_op(_DECODE, coeffs)
