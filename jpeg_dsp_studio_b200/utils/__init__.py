"""Shared utilities of the round trip (mirror of the reference's ``utils`` package,
minus file I/O which is out of scope)."""

from .constants import JPEG_LUMA_Q50, ZIGZAG_ORDER
from .metrics import (Timer, metrics_from_partials, bitrate_from_partials, psnr_from_sse,
                      compute_psnr_ssim, estimate_bitrate_no_entropy, estimate_bitrate_huffman,
                      encode_jfif)
from .preview import PREVIEW_RESOLUTIONS, preview_size, make_preview
from .test_images import (generate_colored_checkerboard, generate_thin_stripes,
                          generate_gradient, generate_text_edges, generate_chroma_stripes,
                          generate_photo, generate_demo_image)

__all__ = [
    'JPEG_LUMA_Q50', 'ZIGZAG_ORDER', 'Timer', 'metrics_from_partials',
    'bitrate_from_partials', 'psnr_from_sse', 'compute_psnr_ssim', 'estimate_bitrate_no_entropy',
    'estimate_bitrate_huffman', 'encode_jfif',
    'generate_colored_checkerboard', 'generate_thin_stripes', 'generate_gradient',
    'generate_text_edges', 'generate_chroma_stripes', 'generate_photo',
    'generate_demo_image', 'PREVIEW_RESOLUTIONS', 'preview_size', 'make_preview',
]
