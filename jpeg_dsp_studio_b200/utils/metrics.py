"""Metric finalisation and timing (mirror of the reference's ``utils/metrics.py``).

The reference computes PSNR/SSIM with scikit-image on the host
(utils/metrics.py:9-28) and estimates the bit count with NumPy (:51-92).  Here the
heavy part - the sums over pixels, windows and coefficients - is done by the CUDA
kernels and arrives as a ``jds_metrics`` record of partials (include/jds.h); the
functions below only turn those partials into the reference's floats with the
reference's own final formulas.  Nothing here touches pixels.
"""

import time
from typing import Dict

import numpy as np

BITRATE_LABEL = 'Estimated (no entropy coding)'


def psnr_from_sse(sse, n_samples: int, data_range: float = 255.0) -> float:
    """skimage.metrics.peak_signal_noise_ratio from a sum of squared errors
    (``10 * log10(data_range**2 / mse)``; identical images give ``inf``)."""
    err = np.float64(sse) / np.float64(n_samples)
    with np.errstate(divide='ignore'):
        return float(10 * np.log10((data_range ** 2) / err))


def metrics_from_partials(m, height: int, width: int) -> Dict[str, float]:
    """``compute_psnr_ssim`` (utils/metrics.py:9-28) from device partials ``m``."""
    n_px = height * width
    out = {
        'psnr_rgb': psnr_from_sse(m.sse_rgb, 3 * n_px),
        'psnr_y': psnr_from_sse(m.sse_y, n_px),
    }
    if m.ssim_count:
        per_ch = np.array([m.ssim_sum[0], m.ssim_sum[1], m.ssim_sum[2]],
                          dtype=np.float64) / np.float64(m.ssim_count)
        out['ssim_rgb'] = float(per_ch.mean())
        out['ssim_y'] = float(np.float64(m.ssim_sum[3]) / np.float64(m.ssim_count))
    else:
        out['ssim_rgb'] = float('nan')
        out['ssim_y'] = float('nan')
    return out


def bitrate_from_partials(m, height: int, width: int) -> Dict:
    """``estimate_bitrate_no_entropy`` (utils/metrics.py:51-92) from the exact integer
    counts of the device: 2 bits per luma-grid block, 6 position bits and
    ``ceil(log2(|v|+1)) + 1`` magnitude bits per non-zero coefficient.

    The reference does this arithmetic in **float32** under NumPy 2 (``np.log2`` of an
    int16 array is float32 and Python ints are weak scalars), so ``bpp`` and
    ``compression_ratio`` carry float32 rounding.  The same float32 operations are applied
    here to the exact integers: bit-identical results whenever the magnitude-bit total is
    below 2**24 (every float32 partial sum of the reference is then exact); above that the
    reference's pairwise float32 summation may differ from the correctly rounded total by
    a few ulp (relative 2e-7)."""
    num_pixels = height * width
    original_bits = num_pixels * 3 * 8
    block_overhead_bits = 2 * int(m.luma_blocks)
    nnz = int(m.nnz)
    exact_bits = block_overhead_bits + int(m.coeff_bits)
    if nnz > 0:
        magnitude_bits = np.float32(int(m.coeff_bits) - 6 * nnz)       # np.sum(... float32 ...)
        coeff_bits = np.float32(6 * nnz) + magnitude_bits              # int (weak) + float32
        estimated = np.float32(block_overhead_bits) + coeff_bits
        bpp = float(estimated / np.float32(num_pixels))
        ratio = float(np.float32(original_bits) / max(estimated, np.float32(1)))
        est_int = int(estimated)
    else:
        estimated = block_overhead_bits                                  # stays a Python int
        bpp = float(estimated / num_pixels)
        ratio = float(original_bits / max(estimated, 1))
        est_int = int(estimated)
    return {
        'estimated_bits': est_int,
        'exact_bits': exact_bits,
        'bpp': bpp,
        'compression_ratio': ratio,
        'nonzero_count': nnz,
        'total_coeffs': int(m.total_coeffs),
        'label': BITRATE_LABEL,
    }


class Timer:
    """Encode/decode stopwatch with the reference's interface (utils/metrics.py:31-48)."""

    def __init__(self):
        self.encode_time_ms = 0.0
        self.decode_time_ms = 0.0

    def measure_encode(self, func, *args, **kwargs):
        start = time.perf_counter()
        result = func(*args, **kwargs)
        self.encode_time_ms = (time.perf_counter() - start) * 1000.0
        return result

    def measure_decode(self, func, *args, **kwargs):
        start = time.perf_counter()
        result = func(*args, **kwargs)
        self.decode_time_ms = (time.perf_counter() - start) * 1000.0
        return result


# ---------------------------------------------------------------------------------------
# the reference's two stand-alone metric functions, computed by the CUDA kernels
# ---------------------------------------------------------------------------------------
def compute_psnr_ssim(original_rgb, reconstructed_rgb) -> Dict[str, float]:
    """PSNR and SSIM on RGB and on BT.601 Y of two uint8 RGB images - the reference's
    ``compute_psnr_ssim`` (utils/metrics.py:9-28).  The squared errors and the 7x7-window
    SSIM sums are reduced on the GPU (jds_compare_images); PSNR_rgb is bit-identical (integer
    sums), PSNR_y within 1e-9 dB, SSIM within 1e-5 of scikit-image's."""
    import ctypes as C
    from .. import _native as N
    from ..engine import get_engine, _is_torch
    a, b = original_rgb, reconstructed_rgb
    if tuple(a.shape) != tuple(b.shape):
        raise ValueError("Input images must have the same dimensions.")      # skimage's message
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError(f"expected H x W x 3 images, got shape {tuple(a.shape)}")
    if min(a.shape[:2]) < 7:
        raise ValueError(
            "win_size exceeds image extent. Either ensure that your images are at least "
            "7x7; or pass win_size explicitly in the function call, with an odd value "
            "less than or equal to the smaller side of your images.")
    eng = get_engine()
    if _is_torch(a) != _is_torch(b):
        raise TypeError("both images must be NumPy arrays or both torch tensors")
    pa, loc, keep_a = eng._in_ptr(a)
    pb, loc_b, keep_b = eng._in_ptr(b)
    if loc != loc_b:
        raise ValueError("both images must live on the same side (host or device)")
    m = N.JdsMetrics()
    h, w = int(a.shape[0]), int(a.shape[1])
    with eng._lock:
        N.check(eng._lib.jds_compare_images(eng._ctx, pa, pb, loc, h, w, C.byref(m)))
    r = metrics_from_partials(m, h, w)
    return {'psnr_rgb': r['psnr_rgb'], 'ssim_rgb': r['ssim_rgb'],
            'psnr_y': r['psnr_y'], 'ssim_y': r['ssim_y']}


class _Partials:
    __slots__ = ("luma_blocks", "nnz", "coeff_bits", "total_coeffs")


def estimate_bitrate_no_entropy(quantized_coeffs, original_shape, block_size: int = 8) -> Dict:
    """Compressed-size estimate without entropy coding - the reference's
    ``estimate_bitrate_no_entropy`` (utils/metrics.py:51-92): 2 bits per block of the luma
    grid, 6 position bits and ``ceil(log2(|v|+1)) + 1`` magnitude bits per non-zero
    coefficient.  The counts come from the GPU (jds_bitrate_partials); the final arithmetic is
    the reference's, including its float32 rounding for int16 input (NumPy 2)."""
    import ctypes as C
    from .. import _native as N
    from ..engine import get_engine
    q = np.asarray(quantized_coeffs)
    if q.dtype.kind not in "iu" or q.dtype.itemsize < 2:
        raise NotImplementedError(f"coefficient dtype {q.dtype}: the pipeline passes int16")
    wide = q.dtype.itemsize > 2                  # np.log2 of int32/int64 is float64: exact sums
    if q.dtype != np.int16:
        if q.size and (q.max() > 32767 or q.min() < -32767):
            raise ValueError("coefficients outside the int16 range")
        q16 = q.astype(np.int16)
    else:
        q16 = q
    q16 = np.ascontiguousarray(q16).reshape(-1)
    h, w = original_shape
    eng = get_engine()
    nnz, bits = C.c_uint64(), C.c_uint64()
    with eng._lock:
        N.check(eng._lib.jds_bitrate_partials(eng._ctx, C.c_void_p(q16.ctypes.data), N.JDS_HOST,
                                              q16.size, C.byref(nnz), C.byref(bits)))
    p = _Partials()
    p.luma_blocks = (-(-h // block_size)) * (-(-w // block_size))
    p.nnz, p.coeff_bits, p.total_coeffs = nnz.value, bits.value, int(q.size)
    if not wide:
        r = bitrate_from_partials(p, h, w)
    else:
        num_pixels, original_bits = h * w, h * w * 24
        overhead = 2 * p.luma_blocks
        if p.nnz:
            estimated = overhead + (6 * p.nnz + np.float64(p.coeff_bits - 6 * p.nnz))
            r = {'estimated_bits': int(estimated), 'bpp': float(estimated / num_pixels),
                 'compression_ratio': float(original_bits / max(estimated, 1))}
        else:
            r = {'estimated_bits': int(overhead), 'bpp': float(overhead / num_pixels),
                 'compression_ratio': float(original_bits / max(overhead, 1))}
        r.update(nonzero_count=p.nnz, total_coeffs=p.total_coeffs, label=BITRATE_LABEL)
    return {k: r[k] for k in ('estimated_bits', 'bpp', 'compression_ratio', 'nonzero_count',
                              'total_coeffs', 'label')}


HUFFMAN_LABEL = 'Baseline JPEG (Huffman, Annex K tables, 3 scans)'
#: bytes of everything but the scans in the file layout the bit count assumes: SOI, APP0/JFIF,
#: one DQT (the reference quantises all three components with the luminance table,
#: engines/pipeline.py:43), SOF0, four DHT segments, three SOS headers, EOI
JFIF_HEADER_BYTES = 2 + 18 + 69 + 19 + (33 + 33 + 183 + 183) + 3 * 10 + 2


def estimate_bitrate_huffman(quantized_coeffs, original_shape, subsampling_mode: str = '4:2:0') -> Dict:
    """What a real baseline JPEG spends on ``all_quantized_coeffs`` - the entropy-coded
    counterpart of ``estimate_bitrate_no_entropy`` (the reference defines
    ``ZIGZAG_ORDER``, utils/constants.py:18-27, but never codes anything).  The three scans
    (Y with the Annex K luminance tables, Cb / Cr with the chrominance tables) are sized
    exactly on the GPU (``jds_entropy_bits``); ``estimated_bits`` adds the byte padding of
    each scan and the fixed headers (0xFF byte stuffing - data dependent, well under 1 % at
    ordinary qualities and a few per cent at very low ones - is not included).  Same keys as ``estimate_bitrate_no_entropy`` plus
    ``scan_bits``."""
    from ..engine import get_engine
    h, w = original_shape
    q = np.asarray(quantized_coeffs)
    if q.dtype != np.int16:
        if q.size and (q.max() > 32767 or q.min() < -32767):
            raise ValueError("coefficients outside the int16 range")
        q = q.astype(np.int16)
    scans = get_engine().entropy_bits(q.reshape(-1), h, w, subsampling_mode)
    payload_bytes = sum((b + 7) // 8 for b in scans)
    total_bits = 8 * (JFIF_HEADER_BYTES + payload_bytes)
    num_pixels = h * w
    return {
        'estimated_bits': int(total_bits),
        'scan_bits': scans,
        'bpp': float(total_bits / num_pixels),
        'compression_ratio': float(num_pixels * 24 / max(total_bits, 1)),
        'nonzero_count': int(np.count_nonzero(q)),
        'total_coeffs': int(q.size),
        'label': HUFFMAN_LABEL,
    }


def encode_jfif(quantized_coeffs, original_shape, subsampling_mode: str = '4:2:0', quality: int = 50,
                qtable=None) -> bytes:
    """The baseline JPEG file the round trip's ``all_quantized_coeffs`` stand for - the
    reference stops at an estimate (utils/metrics.py:57-61) and never uses its
    ``ZIGZAG_ORDER`` (utils/constants.py:18-27).  Zig-zag scan, DC prediction, Annex K Huffman
    coding, padding and 0xFF stuffing run on the GPU (``jds_jfif_encode``); any JPEG decoder
    reads the result.  ``qtable`` defaults to the table of ``quality``
    (engines/quantizer.py:7-19), the one the round trip quantised with."""
    from ..engine import get_engine
    h, w = original_shape
    q = np.asarray(quantized_coeffs)
    if q.dtype != np.int16:
        if q.size and (q.max() > 32767 or q.min() < -32767):
            raise ValueError("coefficients outside the int16 range")
        q = q.astype(np.int16)
    data, _ = get_engine().jfif_encode(q.reshape(-1), h, w, subsampling_mode, quality, qtable)
    return data
