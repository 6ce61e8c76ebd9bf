"""The oracle (oracle/numpy_port.py) against the committed golden vectors that the
UNMODIFIED reference produced (tests/golden/make_golden.py).  CPU only."""

import numpy as np
import pytest

from oracle import numpy_port as P
from tests import cases as C
from tests.conftest import parse_float

SMALL = [c.name for c in C.CASES if not c.big]


@pytest.mark.parametrize("name", SMALL)
def test_port_matches_reference_golden(name, golden_cases):
    case, g = C.BY_NAME[name], golden_cases[name]
    img = case.image()
    assert C.sha(img) == g["input_sha256"], "input generator drifted"
    o = P.compress_reconstruct(img, case.quality, case.mode, case.prefilter, case.sel)
    # bit-exact: integer / byte work
    assert o["all_quantized_coeffs"].dtype == np.int16
    assert o["all_quantized_coeffs"].size == g["coeffs_len"]
    assert C.sha(o["all_quantized_coeffs"]) == g["coeffs_sha256"]
    assert C.sha(o["reconstructed_image"]) == g["recon_sha256"]
    assert C.sha(o["error_map_y"]) == g["error_map_y_sha256"]
    assert C.sha(o["error_map_rgb"]) == g["error_map_rgb_sha256"]
    assert [int(v) for v in o["quantized_histogram"]] == g["histogram"]
    assert o["nonzero_coeffs"] == g["nonzero_coeffs"]
    assert o["total_coeffs"] == g["total_coeffs"]
    # the reference's own float32-accumulated bit estimate, reproduced exactly
    assert o["bpp"] == g["bpp"]
    assert o["compression_ratio"] == g["compression_ratio"]
    for k in ("psnr_y", "psnr_rgb", "ssim_y", "ssim_rgb"):
        assert o[k] == parse_float(g[k]), k
    if g["selected"] is None:
        assert o["selected_block"] is None
    else:
        for k, h in g["selected"].items():
            assert C.sha(o["selected_block"][k]) == h, k


def test_port_matches_reference_sweep(golden_sweep):
    """cfg4 at test size: every 7th quality plus the table extremes, all 4 combos."""
    name, make = C.SWEEP_IMAGE
    img = make()
    assert C.sha(img) == golden_sweep["input_sha256"]
    pick = set(range(1, 101, 7)) | {1, 2, 49, 50, 51, 99, 100}
    n = 0
    for pt in golden_sweep["points"]:
        if pt["quality"] not in pick:
            continue
        o = P.compress_reconstruct(img, pt["quality"], pt["mode"], pt["prefilter"],
                                   want_maps=False)
        assert C.sha(o["all_quantized_coeffs"]) == pt["coeffs_sha256"], pt
        assert C.sha(o["reconstructed_image"]) == pt["recon_sha256"], pt
        assert o["bpp"] == pt["bpp"]
        assert o["psnr_y"] == parse_float(pt["psnr_y"])
        assert o["ssim_rgb"] == pt["ssim_rgb"]
        n += 1
    assert n == len(pick) * 4


def test_quant_table_extremes():
    """Q=1 -> all 255, Q=50 -> base table, Q=100 -> all ones."""
    assert np.all(P.scale_quant_matrix(1) == 255)
    assert np.array_equal(P.scale_quant_matrix(50), P.JPEG_LUMA_Q50)
    assert np.all(P.scale_quant_matrix(100) == 1)


def test_fma_emulation_against_libm():
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    rng = np.random.default_rng(1)
    a = rng.uniform(-300, 300, 20000)
    b = rng.uniform(0, 1, 20000)
    c = rng.uniform(-300, 300, 20000)
    ref = np.array([libm.fma(x, y, z) for x, y, z in zip(a, b, c)])
    assert np.array_equal(P._fma(a, b, c), ref)
    assert np.any(ref != a * b + c)          # the test would be vacuous otherwise
