#!/usr/bin/env python
"""Generate tests/golden/preview.json: SHA-256 of the preview image the reference's
_update_preview_image (gui/compression_tab.py:532-552) produces - its own size arithmetic and
live cv2.resize(INTER_AREA) - for the cases of tests/test_preview_cpu.py.  Build container only.

    python tests/golden/make_preview_golden.py
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.test_preview_cpu import PREVIEW_CASES, _img  # noqa: E402


def reference_preview(image, target_w, target_h):
    # gui/compression_tab.py:538-552, statement for statement (the class needs PySide6)
    h, w = image.shape[:2]
    if w <= target_w and h <= target_h:
        return image.copy()
    scale = min(target_w / w, target_h / h)
    new_w = int(w * scale)
    new_h = int(h * scale)
    return cv2.resize(image, (new_w, new_h), interpolation=cv2.INTER_AREA)


out = []
for name, shape, seed, target in PREVIEW_CASES:
    img = _img(shape, seed)
    pv = reference_preview(img, *target)
    out.append({"name": name, "input_shape": list(img.shape), "target": list(target),
                "input_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
                "shape": list(pv.shape),
                "sha256": hashlib.sha256(np.ascontiguousarray(pv).tobytes()).hexdigest()})
    print(name, pv.shape)
json.dump({"meta": {"generator": "tests/golden/make_preview_golden.py", "cv2": cv2.__version__,
                    "cv2_ipp": bool(cv2.ipp.useIPP()), "numpy": np.__version__},
           "cases": out}, open(os.path.join(HERE, "preview.json"), "w"), indent=1)
