"""Stand-in for ``skimage.metrics`` (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference calls ``peak_signal_noise_ratio`` and ``structural_similarity`` at
``/root/reference/utils/metrics.py:11-14,20-21``.  scikit-image is not installed in
this image and cannot be fetched, so these follow its published algorithm
(skimage >= 0.21, ``skimage/metrics/simple_metrics.py`` and
``skimage/metrics/_structural_similarity.py``) for exactly the call shapes the
reference uses: ``data_range=255``, default 7x7 uniform window, K1=0.01, K2=0.03,
sample covariance, optional ``channel_axis``.
"""

import numpy as np
from scipy.ndimage import uniform_filter

__all__ = ["peak_signal_noise_ratio", "structural_similarity", "install"]


def _as_f64(a):
    return np.asarray(a).astype(np.float64)


def peak_signal_noise_ratio(image_true, image_test, *, data_range=None):
    if data_range is None:
        raise ValueError("stand-in needs data_range (the reference always passes 255)")
    a, b = _as_f64(image_true), _as_f64(image_test)
    if a.shape != b.shape:
        raise ValueError("Input images must have the same dimensions.")
    err = np.mean((a - b) ** 2, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return 10 * np.log10((data_range ** 2) / err)


def _ssim_2d(x, y, data_range, win_size=7, K1=0.01, K2=0.03):
    if min(x.shape) < win_size:
        raise ValueError(
            "win_size exceeds image extent. Either ensure that your images are at "
            "least 7x7; or pass win_size explicitly in the function call, with an "
            "odd value less than or equal to the smaller side of your images.")
    NP = win_size ** x.ndim
    cov_norm = NP / (NP - 1)
    ux = uniform_filter(x, size=win_size)
    uy = uniform_filter(y, size=win_size)
    uxx = uniform_filter(x * x, size=win_size)
    uyy = uniform_filter(y * y, size=win_size)
    uxy = uniform_filter(x * y, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1 = (K1 * data_range) ** 2
    C2 = (K2 * data_range) ** 2
    A1, A2, B1, B2 = (2 * ux * uy + C1, 2 * vxy + C2,
                      ux ** 2 + uy ** 2 + C1, vx + vy + C2)
    S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    return S[pad:S.shape[0] - pad, pad:S.shape[1] - pad].mean(dtype=np.float64)


def structural_similarity(im1, im2, *, data_range=None, channel_axis=None, **kw):
    if kw:
        raise TypeError(f"stand-in does not implement {sorted(kw)}")
    if data_range is None:
        raise ValueError("stand-in needs data_range (the reference always passes 255)")
    a, b = _as_f64(im1), _as_f64(im2)
    if a.shape != b.shape:
        raise ValueError("Input images must have the same dimensions.")
    if channel_axis is None:
        return _ssim_2d(a, b, data_range)
    a = np.moveaxis(a, channel_axis, -1)
    b = np.moveaxis(b, channel_axis, -1)
    per_ch = np.empty(a.shape[-1], dtype=np.float64)
    for c in range(a.shape[-1]):
        per_ch[c] = _ssim_2d(np.ascontiguousarray(a[..., c]),
                             np.ascontiguousarray(b[..., c]), data_range)
    return per_ch.mean()


def install():
    """Register this module as ``skimage.metrics`` so the reference imports."""
    import sys
    import types
    if "skimage.metrics" in sys.modules:
        return
    pkg = types.ModuleType("skimage")
    pkg.__path__ = []
    met = types.ModuleType("skimage.metrics")
    met.peak_signal_noise_ratio = peak_signal_noise_ratio
    met.structural_similarity = structural_similarity
    pkg.metrics = met
    sys.modules["skimage"] = pkg
    sys.modules["skimage.metrics"] = met
