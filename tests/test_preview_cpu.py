"""SURVEY 8f #3 - the GUI's preview downscale (gui/compression_tab.py:532-552).  CPU side:
the oracle's restatement of cv2.resize(uint8, INTER_AREA) against live OpenCV (the third-party
library the reference calls; present in this image) and against committed golden digests, and
the size rule against the reference's Python arithmetic."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import numpy_port as P
from tests import cases as CS

GOLD = os.path.join(os.path.dirname(__file__), "golden", "preview.json")

PREVIEW_CASES = [
    # name, (H, W), seed, target (w, h)
    ("4k_to_1080p", (2160, 3840), 31, (1920, 1080)),          # 2x2 integer path
    ("4k_to_720p", (2160, 3840), 31, (1280, 720)),            # 3x3 integer path
    ("4k_to_540p", (2160, 3840), 31, (960, 540)),             # 4x4
    ("1080p_to_720p", (1080, 1920), 32, (1280, 720)),         # 1.5: general path
    ("1440p_to_1080p", (1440, 2560), 33, (1920, 1080)),
    ("photo_3000x4000_to_720p", (3000, 4000), 34, (1280, 720)),   # limited by height
    ("tall_1500x1000_to_540p", (1500, 1000), 35, (960, 540)),
    ("odd_777x1234_to_540p", (777, 1234), 36, (960, 540)),
    ("fits_600x800", (600, 800), 37, (1280, 720)),            # kept as is
]


def _img(shape, seed):
    return CS.rand_rgb(seed, shape[0], shape[1])


def test_preview_size_rule():
    assert P.preview_size(2160, 3840, 1280, 720) == (720, 1280)
    assert P.preview_size(600, 800, 1280, 720) == (600, 800)
    assert P.preview_size(3000, 4000, 1280, 720) == (720, 960)
    assert P.preview_size(1500, 1000, 960, 540) == (540, 360)
    assert P.preview_size(777, 1234, 960, 540) == (540, int(1234 * min(960 / 1234, 540 / 777)))


def test_preview_size_c_abi_matches_python_arithmetic():
    from jpeg_dsp_studio_b200.utils.preview import preview_size
    rng = np.random.default_rng(0)
    for _ in range(2000):
        h, w = int(rng.integers(1, 9000)), int(rng.integers(1, 9000))
        for tw, th, _label in ((960, 540, 0), (1280, 720, 0), (1920, 1080, 0)):
            assert preview_size(h, w, (tw, th)) == P.preview_size(h, w, tw, th)


@pytest.mark.parametrize("name,shape,seed,target", PREVIEW_CASES, ids=[c[0] for c in PREVIEW_CASES])
def test_oracle_resize_matches_golden(name, shape, seed, target):
    gold = {r["name"]: r for r in json.load(open(GOLD))["cases"]}[name]
    img = _img(shape, seed)
    assert hashlib.sha256(img.tobytes()).hexdigest() == gold["input_sha256"]
    out = P.make_preview(img, *target)
    assert list(out.shape) == gold["shape"]
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == gold["sha256"]


def test_oracle_resize_matches_live_opencv_on_random_sizes():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for _ in range(40):
        h, w = int(rng.integers(8, 700)), int(rng.integers(8, 700))
        dh, dw = int(rng.integers(1, h + 1)), int(rng.integers(1, w + 1))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA)
        assert np.array_equal(P.resize_area_u8(img, dh, dw), ref), (h, w, dh, dw)
