"""Colour conversion and chroma resampling - the reference's ``engines/color_space.py`` API
with every function a CUDA operator of libjds.so (csrc/jds_ops.cu), exact fp64 arithmetic in
the reference's operation order: bit-identical to NumPy / OpenCV as the reference calls them
(SURVEY Appendix A1-A3, A8, A9).  Inside ``compress_reconstruct`` the same arithmetic runs
fused; these stand-alone forms exist for callers (and tests) that use a stage on its own."""

import ctypes as C
from typing import Tuple

import numpy as np

from .. import _native as N
from ..engine import get_engine

_MODES = {'4:4:4': N.JDS_SUB_444, '4:2:2': N.JDS_SUB_422, '4:2:0': N.JDS_SUB_420}


def _convert(direction: int, img: np.ndarray) -> np.ndarray:
    if img.ndim < 3:
        # the reference indexes rgb[:, :, 0] (engines/color_space.py:10)
        raise IndexError(f"too many indices for array: array is {img.ndim}-dimensional, "
                         "but 3 were indexed")
    a = np.ascontiguousarray(img[:, :, :3], dtype=np.float64)
    out = np.empty_like(a)
    if a.size == 0:
        return out
    eng = get_engine()
    with eng._lock:
        N.check(eng._lib.jds_color_convert(eng._ctx, direction, a.size // 3,
                                           C.c_void_p(a.ctypes.data), C.c_void_p(out.ctypes.data)))
    return out


def rgb_to_ycbcr(rgb: np.ndarray) -> np.ndarray:
    """RGB to YCbCr, ITU-R BT.601 full range (engines/color_space.py:8-14)."""
    return _convert(0, rgb)


def ycbcr_to_rgb(ycbcr: np.ndarray) -> np.ndarray:
    """YCbCr to RGB, clipped to [0, 255] (engines/color_space.py:17-24)."""
    return _convert(1, ycbcr)


def _subsample_plane(eng, plane: np.ndarray, code: int, prefilter: bool) -> np.ndarray:
    p = np.ascontiguousarray(plane, dtype=np.float64)
    h, w = p.shape
    ch, cw = C.c_int(), C.c_int()
    N.check(eng._lib.jds_plane_dims(h, w, code, C.byref(ch), C.byref(cw)))
    out = np.empty((ch.value, cw.value), dtype=np.float64)
    with eng._lock:
        N.check(eng._lib.jds_subsample_plane(eng._ctx, C.c_void_p(p.ctypes.data), h, w, code,
                                             int(bool(prefilter)), C.c_void_p(out.ctypes.data)))
    return out


def subsample_chroma(cb: np.ndarray, cr: np.ndarray, mode: str,
                     use_prefilter: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """Chroma decimation (engines/color_space.py:27-53): optional 3x3 Gaussian prefilter
    (sigma 0.75), then area averaging to W//2 (4:2:2) or W//2 x H//2 (4:2:0)."""
    if mode == '4:4:4':
        return cb.copy(), cr.copy()
    if mode not in _MODES:
        raise ValueError(f"Unknown subsampling mode: {mode}")
    eng = get_engine()
    return (_subsample_plane(eng, cb, _MODES[mode], use_prefilter),
            _subsample_plane(eng, cr, _MODES[mode], use_prefilter))


def _upsample_plane(eng, plane: np.ndarray, shape: Tuple[int, int]) -> np.ndarray:
    p = np.ascontiguousarray(plane, dtype=np.float64)
    h, w = p.shape
    out = np.empty((int(shape[0]), int(shape[1])), dtype=np.float64)
    with eng._lock:
        N.check(eng._lib.jds_upsample_plane(eng._ctx, C.c_void_p(p.ctypes.data), h, w,
                                            C.c_void_p(out.ctypes.data), out.shape[0], out.shape[1]))
    return out


def upsample_chroma(cb_sub: np.ndarray, cr_sub: np.ndarray, target_shape: Tuple[int, int],
                    method: str = 'bilinear') -> Tuple[np.ndarray, np.ndarray]:
    """Chroma upsampling to ``target_shape`` = (H, W) (engines/color_space.py:56-66).  The
    pipeline only ever uses 'bilinear' (engines/pipeline.py:90)."""
    if method != 'bilinear':
        raise NotImplementedError("only method='bilinear' (the one the pipeline uses) is provided")
    eng = get_engine()
    return _upsample_plane(eng, cb_sub, target_shape), _upsample_plane(eng, cr_sub, target_shape)
