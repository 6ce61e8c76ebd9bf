"""Compression parameters - same fields, defaults and validation as the reference's
``models/compression_params.py:7-20`` (the config surface of the drop-in)."""

from dataclasses import dataclass
from typing import Literal


@dataclass
class CompressionParams:
    """JPEG-like compression parameters (reference: models/compression_params.py:7-20)."""

    block_size: int = 8
    quality: int = 50
    subsampling_mode: Literal['4:4:4', '4:2:2', '4:2:0'] = '4:2:0'
    use_prefilter: bool = False

    def __post_init__(self):
        # Same checks, same messages, same order as the reference (:16-20).  The
        # subsampling mode is deliberately NOT validated here: the reference raises
        # for an unknown mode only inside subsample_chroma (color_space.py:51).
        if not (1 <= self.quality <= 100):
            raise ValueError(f"Quality must be 1-100, got {self.quality}")
        if self.block_size not in [4, 8, 16, 32]:
            raise ValueError(f"Block size must be 4, 8, 16, or 32, got {self.block_size}")
