"""Preview mode of the reference GUI (gui/compression_tab.py:32-36, 532-558): large images
are shrunk with cv2.resize(INTER_AREA) before they enter the round trip.  Here the resize is
a CUDA kernel of libjds.so (csrc/jds_preview.cu), bit-identical to OpenCV's."""

import ctypes as C
from typing import Optional, Tuple

from .. import _native as N

# gui/compression_tab.py:32-36
PREVIEW_RESOLUTIONS = [
    (960, 540, "960×540 (Fast)"),
    (1280, 720, "1280×720 (Balanced)"),
    (1920, 1080, "1920×1080 (High)"),
]
DEFAULT_PREVIEW = (1280, 720)            # combo index 1 (gui/compression_tab.py:218)


def preview_size(height: int, width: int, target: Tuple[int, int] = DEFAULT_PREVIEW) -> Tuple[int, int]:
    """(new_h, new_w) of _update_preview_image (gui/compression_tab.py:538-547)."""
    oh, ow = C.c_int(), C.c_int()
    N.check(N.load().jds_preview_size(int(height), int(width), int(target[0]), int(target[1]),
                                      C.byref(oh), C.byref(ow)))
    return oh.value, ow.value


def make_preview(image, target: Tuple[int, int] = DEFAULT_PREVIEW, *, device: Optional[int] = None):
    """The image the reference would process in preview mode: a copy when it already fits
    ``target`` = (width, height), otherwise the INTER_AREA downscale
    (gui/compression_tab.py:532-552).  NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out."""
    from ..engine import get_engine
    h, w = int(image.shape[0]), int(image.shape[1])
    nh, nw = preview_size(h, w, target)
    if (nh, nw) == (h, w):
        return image.clone() if hasattr(image, "clone") else image.copy()
    return get_engine(device).resize_area(image, nh, nw)
