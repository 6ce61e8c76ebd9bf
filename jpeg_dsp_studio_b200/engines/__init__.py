"""DSP engines of the round trip (mirror of the reference's ``engines`` package)."""

from ..utils.constants import JPEG_LUMA_Q50
from .pipeline import compress_reconstruct, quality_sweep, compress_batch

__all__ = ['JPEG_LUMA_Q50', 'compress_reconstruct', 'quality_sweep', 'compress_batch']
