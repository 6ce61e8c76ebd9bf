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
