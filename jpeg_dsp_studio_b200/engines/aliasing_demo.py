"""Chroma-aliasing demo (mirror of the reference's ``gui/dialogs/aliasing_demo_dialog.py``
worker) - the computational part of the dialog, without Qt:

  * the three synthetic patterns                               (aliasing_demo_dialog.py:20-66)
  * ``compute_metrics(original, reconstructed)``               (:69-83)
  * ``AliasingDemoWorker(image, quality).run()``               (:86-166)

The reference runs OpenCV *float32* kernels on the host (YCrCb conversion, 5x5 Gaussian,
``[::2, ::2]``, bilinear resize) and then its NumPy hot path at 4:4:4.  Here both arms run on
the GPU (``jds_aliasing_demo``: csrc/jds_alias.cu + the round-trip kernels) and return arrays
bit-identical to the reference's in the default ``precision='exact'``.
"""

from typing import Callable, Dict, Optional

import numpy as np

from ..engine import get_engine
from ..models import CompressionParams


def generate_equiluminance_stripes(size: int = 256) -> np.ndarray:
    """Red/cyan 1-pixel vertical stripes with matched luminance (:20-34)."""
    img = np.empty((size, size, 3), dtype=np.uint8)
    img[:, 0::2] = np.array([220, 40, 60], dtype=np.uint8)
    img[:, 1::2] = np.array([30, 220, 210], dtype=np.uint8)
    return img


def _two_colour(mask: np.ndarray, a, b) -> np.ndarray:
    return np.where(mask[..., None], np.array(a, dtype=np.uint8), np.array(b, dtype=np.uint8))


def generate_chroma_checkerboard(size: int = 256) -> np.ndarray:
    """2x2-pixel checkerboard with high chroma contrast (:37-50)."""
    i, j = np.indices((size, size))
    return _two_colour(((i // 2) + (j // 2)) % 2 == 0, [230, 50, 60], [50, 220, 220])


def generate_1px_checkerboard(size: int = 256) -> np.ndarray:
    """1x1-pixel checkerboard - maximum spatial frequency (:53-66)."""
    i, j = np.indices((size, size))
    return _two_colour((i + j) % 2 == 0, [240, 40, 50], [40, 240, 230])


def compute_metrics(original: np.ndarray, reconstructed: np.ndarray, *, device: Optional[int] = None) -> Dict[str, float]:
    """PSNR / SSIM of the RGB frames and of OpenCV's integer luma (:69-83), reduced on the GPU."""
    return get_engine(device).aliasing_metrics(original, reconstructed)


class AliasingDemoWorker:
    """The reference's worker without the Qt signals: ``run()`` returns the dict the reference
    emits through ``finished`` (original, recon_no_pf, recon_pf, diff_no_pf, diff_pf,
    metrics_no_pf, metrics_pf); ``progress`` (optional callable) receives the two messages."""

    def __init__(self, image: np.ndarray, quality: int = 50, *, device: Optional[int] = None,
                 precision: str = 'exact', progress: Optional[Callable[[str], None]] = None):
        self.image = image
        self.quality = quality
        self._device, self._precision, self._progress = device, precision, progress

    def _process_with_explicit_subsample(self, prefilter: bool) -> np.ndarray:
        return self._arm(prefilter)['recon']

    def _arm(self, prefilter: bool) -> dict:
        CompressionParams(quality=self.quality, block_size=8, subsampling_mode='4:4:4', use_prefilter=False)
        return get_engine(self._device).aliasing_demo_arm(self.image, self.quality, prefilter,
                                                          precision=self._precision)

    def run(self) -> dict:
        say = self._progress or (lambda msg: None)
        say("Processing without prefilter (true decimation)...")
        a = self._arm(False)
        say("Processing with Gaussian prefilter...")
        b = self._arm(True)
        return {
            'original': self.image,
            'recon_no_pf': a['recon'], 'recon_pf': b['recon'],
            'diff_no_pf': a['diff'], 'diff_pf': b['diff'],
            'metrics_no_pf': a['metrics'], 'metrics_pf': b['metrics'],
        }
