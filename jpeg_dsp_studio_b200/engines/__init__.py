"""DSP engines of the round trip (mirror of the reference's ``engines`` package:
engines/__init__.py:3-27).  Inside ``compress_reconstruct`` colour conversion, chroma resampling and block
splitting are fused into the kernels; every stage is also exported here as a stand-alone
operator with the reference's name and signature."""

from ..utils.constants import JPEG_LUMA_Q50
from .color_space import rgb_to_ycbcr, ycbcr_to_rgb, subsample_chroma, upsample_chroma
from .block_processor import pad_to_multiple, split_into_blocks, merge_blocks
from .dct_engine import dct2, idct2, encode_block, decode_block
from .quantizer import scale_quant_matrix, quantize, dequantize
from .pipeline import compress_reconstruct, quality_sweep, compress_batch, compress_stream, plot_payload

__all__ = ['rgb_to_ycbcr', 'ycbcr_to_rgb', 'subsample_chroma', 'upsample_chroma', 'pad_to_multiple',
           'split_into_blocks', 'merge_blocks', 'dct2', 'idct2', 'encode_block', 'decode_block', 'scale_quant_matrix', 'quantize',
           'dequantize', 'JPEG_LUMA_Q50', 'compress_reconstruct', 'quality_sweep',
           'compress_batch', 'compress_stream', 'plot_payload']
