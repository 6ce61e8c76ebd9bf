"""SURVEY 8f #3 on the GPU: the preview downscale kernels (csrc/jds_preview.cu) against the
golden digests of cv2.resize(INTER_AREA) and against the oracle on random shapes; then the
GUI's preview flow - downscale, round trip on the (often odd-sized) preview."""
import hashlib
import json

import numpy as np
import pytest

from tests import cases as CS
from tests.test_preview_cpu import GOLD, PREVIEW_CASES, _img

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import jpeg_dsp_studio_b200 as J
    return J


@pytest.fixture(scope="module")
def oracle():
    from oracle import numpy_port
    return numpy_port


@pytest.mark.parametrize("name,shape,seed,target", PREVIEW_CASES, ids=[c[0] for c in PREVIEW_CASES])
def test_preview_matches_opencv_golden(J, name, shape, seed, target):
    from jpeg_dsp_studio_b200.utils.preview import make_preview
    gold = {r["name"]: r for r in json.load(open(GOLD))["cases"]}[name]
    img = _img(shape, seed)
    out = make_preview(img, target)
    assert out is not img and list(out.shape) == gold["shape"] and out.dtype == np.uint8
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == gold["sha256"]


def test_resize_area_random_shapes_against_oracle(J, oracle):
    eng = J.get_engine()
    rng = np.random.default_rng(5)
    for _ in range(30):
        h, w = int(rng.integers(8, 900)), int(rng.integers(8, 900))
        dh, dw = int(rng.integers(1, h + 1)), int(rng.integers(1, w + 1))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(eng.resize_area(img, dh, dw), oracle.resize_area_u8(img, dh, dw)), (h, w, dh, dw)
    # extreme factors: one output pixel, identity
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    assert np.array_equal(eng.resize_area(img, 1, 1), oracle.resize_area_u8(img, 1, 1))
    assert np.array_equal(eng.resize_area(img, 37, 53), img)
    with pytest.raises(Exception):
        eng.resize_area(img, 40, 53)                     # enlarging is not the preview path


def test_preview_device_tensor_then_round_trip(J, oracle):
    """GUI flow in preview mode (gui/compression_tab.py:554-575): the round trip runs on the
    preview; the whole chain stays on the device and matches the oracle run on OpenCV's
    preview."""
    import torch
    from jpeg_dsp_studio_b200.utils.preview import make_preview
    img = CS.photo_tiled(1301, 1951)                      # -> 720 x 1079: odd width, 4:2:0
    pv_ref = oracle.make_preview(img, 1280, 720)
    assert pv_ref.shape == (720, 1079, 3)
    pv = make_preview(torch.from_numpy(img).cuda(), (1280, 720))
    assert pv.is_cuda and np.array_equal(pv.cpu().numpy(), pv_ref)
    o = J.get_engine().roundtrip(pv, 50, "4:2:0", False, precision="exact", want_coeffs=True)
    ref = oracle.compress_reconstruct(pv_ref, 50, "4:2:0", False, want_maps=False)
    assert np.array_equal(o.coeffs.cpu().numpy(), ref["all_quantized_coeffs"])
    assert np.array_equal(o.recon.cpu().numpy(), ref["reconstructed_image"])
    res, _ = J.compress_reconstruct(pv_ref, J.CompressionParams(quality=50))
    assert res.psnr_rgb == ref["psnr_rgb"] and res.bpp == ref["bpp"]
