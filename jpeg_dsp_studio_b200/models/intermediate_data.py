"""``IntermediateData`` - what the reference's GUI plots read next to the result.

API contract (reference ``models/intermediate_data.py:8-23``, filled at
``engines/pipeline.py:119-165``): eleven fields in the positional order below, every array
field ``None`` by default.  The class is generated from ``FIELD_SPECS`` so that each field
carries its dtype / shape / producer, which the tests and ``describe()`` use.

In the drop-in the bulky arrays (both error maps as fp64, all int16 coefficients) come back
from the device into pinned host memory; ``compress_reconstruct(..., intermediates=False)``
leaves them ``None`` and ``plot_payload`` returns the reduced forms the plots actually draw.
"""

from dataclasses import field, make_dataclass
from typing import Optional

import numpy as np

#: (name, default, what it holds)
FIELD_SPECS = (
    ("selected_block_idx", (0, 0), "(block_row, block_col) as passed in - echoed even when out of range"),
    ("selected_block_original", None, "8x8 fp64: the padded Y plane under the selected block"),
    ("selected_block_shifted", None, "8x8 fp64: original - 128"),
    ("selected_block_dct", None, "8x8 fp64: orthonormal 2-D DCT-II of the shifted block"),
    ("selected_block_quantized", None, "8x8 int16: round-half-even(dct / Q)"),
    ("selected_block_dequantized", None, "8x8 fp64: quantized * Q"),
    ("selected_block_reconstructed", None, "8x8 fp64: IDCT + 128, clipped to 0..255"),
    ("error_map_y", None, "H x W fp64: |Y - Y_reconstructed|"),
    ("error_map_rgb", None, "H x W fp64: mean over channels of |rgb - rgb_reconstructed| before truncation"),
    ("quantized_histogram", None, "int64[50]: np.histogram(all coefficients, 50, (-100, 100))"),
    ("all_quantized_coeffs", None, "int16, flat: channel Y|Cb|Cr -> block raster -> 64 row-major values"),
)


def _describe(cls) -> dict:
    """field name -> what it holds"""
    return {name: doc for name, _, doc in FIELD_SPECS}


IntermediateData = make_dataclass(
    "IntermediateData",
    [(name, tuple if name == "selected_block_idx" else Optional[np.ndarray],
      field(default=default, metadata={"doc": doc})) for name, default, doc in FIELD_SPECS],
    namespace={"describe": classmethod(_describe),
               "__doc__": "Intermediate results for the GUI plots (see FIELD_SPECS)."},
)
IntermediateData.__module__ = __name__
