"""DSP engines of the round trip (mirror of the reference's ``engines`` package:
engines/__init__.py:3-27).  Colour conversion, chroma resampling and block splitting exist
only inside the fused kernels; the stand-alone operators below are the ones the reference's
own unit tests exercise."""

from ..utils.constants import JPEG_LUMA_Q50
from .dct_engine import dct2, idct2, encode_block, decode_block
from .quantizer import scale_quant_matrix, quantize, dequantize
from .pipeline import compress_reconstruct, quality_sweep, compress_batch, plot_payload

__all__ = ['dct2', 'idct2', 'encode_block', 'decode_block', 'scale_quant_matrix', 'quantize',
           'dequantize', 'JPEG_LUMA_Q50', 'compress_reconstruct', 'quality_sweep',
           'compress_batch', 'plot_payload']
