"""API types of the round trip (mirror of the reference's ``models`` package)."""

from .compression_params import CompressionParams
from .compression_result import CompressionResult
from .intermediate_data import IntermediateData

__all__ = ['CompressionParams', 'CompressionResult', 'IntermediateData']
