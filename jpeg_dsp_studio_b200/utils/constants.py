"""JPEG tables (reference: utils/constants.py:6-27)."""

import numpy as np

#: Standard JPEG luminance quantisation table, quality 50 (Annex K.1).  The reference
#: uses this one table for Y, Cb and Cr alike (engines/pipeline.py:43).
JPEG_LUMA_Q50 = np.array([
    [16, 11, 10, 16, 24, 40, 51, 61],
    [12, 12, 14, 19, 26, 58, 60, 55],
    [14, 13, 16, 24, 40, 57, 69, 56],
    [14, 17, 22, 29, 51, 87, 80, 62],
    [18, 22, 37, 56, 68, 109, 103, 77],
    [24, 35, 55, 64, 81, 104, 113, 92],
    [49, 64, 78, 87, 103, 121, 120, 101],
    [72, 92, 95, 98, 112, 100, 103, 99]
], dtype=np.float64)


def _zigzag(n: int = 8) -> np.ndarray:
    """ZIGZAG_ORDER[r, c] = raster index of the (8r+c)-th coefficient in zig-zag scan."""
    order = sorted(((i, j) for i in range(n) for j in range(n)),
                   key=lambda p: (p[0] + p[1], p[0] if (p[0] + p[1]) % 2 else p[1]))
    return np.array([i * n + j for i, j in order], dtype=np.int32).reshape(n, n)


#: Zig-zag scan order (defined by the reference, unused by the round trip).
ZIGZAG_ORDER = _zigzag()
