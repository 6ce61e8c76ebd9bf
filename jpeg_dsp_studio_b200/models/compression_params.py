"""``CompressionParams`` - the configuration surface of the drop-in.

API contract (reference ``models/compression_params.py:7-20``): four fields in this
positional order with these defaults, ``ValueError`` with the reference's messages for a
quality outside 1..100 or a block size outside {4, 8, 16, 32}; quality is checked first.
``subsampling_mode`` is NOT validated at construction - the reference raises for an unknown
mode only when the chroma planes are decimated (``engines/color_space.py:51``), and so does
the drop-in (``Engine`` maps the string to ``JDS_SUB_*``).  Block sizes other than 8 pass
validation here and fail inside the round trip, as in the reference.
"""

from dataclasses import dataclass, field, fields
from typing import ClassVar, Literal, Tuple

SubsamplingMode = Literal['4:4:4', '4:2:2', '4:2:0']


@dataclass
class CompressionParams:
    """Parameters of one compression round trip."""

    #: values ``__post_init__`` accepts
    QUALITY_RANGE: ClassVar[Tuple[int, int]] = (1, 100)
    BLOCK_SIZES: ClassVar[Tuple[int, ...]] = (4, 8, 16, 32)

    block_size: int = field(default=8, metadata={"doc": "transform block edge; only 8 runs"})
    quality: int = field(default=50, metadata={"doc": "IJG quality factor, 1..100"})
    subsampling_mode: SubsamplingMode = field(default='4:2:0', metadata={"doc": "chroma layout"})
    use_prefilter: bool = field(default=False, metadata={"doc": "3x3 Gaussian before decimation"})

    def __post_init__(self):
        lo, hi = self.QUALITY_RANGE
        if self.quality < lo or self.quality > hi:
            raise ValueError(f"Quality must be {lo}-{hi}, got {self.quality}")
        if self.block_size not in self.BLOCK_SIZES:
            allowed = ", ".join(str(b) for b in self.BLOCK_SIZES[:-1]) + f", or {self.BLOCK_SIZES[-1]}"
            raise ValueError(f"Block size must be {allowed}, got {self.block_size}")

    @classmethod
    def describe(cls) -> dict:
        """field name -> one-line description (for CLIs / logs)"""
        return {f.name: f.metadata.get("doc", "") for f in fields(cls)}
