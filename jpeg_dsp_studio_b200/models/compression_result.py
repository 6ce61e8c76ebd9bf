"""Round-trip result - field-for-field the reference's ``models/compression_result.py:7-30``."""

from dataclasses import dataclass

import numpy as np


@dataclass
class CompressionResult:
    """Results of one compress/reconstruct call (reference: models/compression_result.py)."""

    original_image: np.ndarray          # the SAME object that was passed in (pipeline.py:103)
    reconstructed_image: np.ndarray     # uint8 H x W x 3

    psnr_y: float
    ssim_y: float
    psnr_rgb: float
    ssim_rgb: float

    bpp: float
    compression_ratio: float
    nonzero_coeffs: int
    total_coeffs: int

    encode_time_ms: float
    decode_time_ms: float

    bitrate_label: str = "Estimated (no entropy coding)"
