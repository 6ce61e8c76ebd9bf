"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol
include/jds.h declares, host-only entry points agree with the oracle, the API
dataclasses validate like the reference's, and compute entry points fail loudly
without a GPU (no fallback).  No kernels are launched here."""

import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import numpy_port as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_jds_build", os.path.join(ROOT, "jpeg_dsp_studio_b200", "build.py"))
    build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(build)
    build.build_native()
    from jpeg_dsp_studio_b200 import _native
    _native.load()
    return _native


def declared_functions():
    text = open(os.path.join(ROOT, "include", "jds.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jds_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native):
    lib = C.CDLL(native.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"libjds.so lacks {n} declared in include/jds.h"
    assert set(names) == set(native.PROTOTYPES), "binding and header disagree"
    assert lib.jds_abi_version() == native.JDS_ABI_VERSION


def test_struct_layouts_match_header(native):
    assert C.sizeof(native.JdsParams) == 32
    assert C.sizeof(native.JdsMetrics) == 8 * (1 + 1 + 4 + 1 + 1 + 1 + 1 + 1 + 50 + 1 + 3)


def test_quant_table_matches_reference_formula(native):
    lib = native.load()
    for q in range(1, 101):
        t = (C.c_double * 64)()
        native.check(lib.jds_quant_table(q, t))
        assert np.array_equal(np.array(t).reshape(8, 8), P.scale_quant_matrix(q)), q
    t = (C.c_double * 64)()
    assert lib.jds_quant_table(0, t) == native.JDS_ERR_INVALID
    assert b"Quality must be 1-100" in lib.jds_last_error()


@pytest.mark.parametrize("h,w,mode,want", [
    (512, 512, "4:2:0", 393216), (1080, 1920, "4:4:4", 6220800),
    (1080, 1920, "4:2:2", 4147200), (2160, 3840, "4:2:0", 12441600),
    (250, 334, "4:2:0", 64 * (32 * 42 + 2 * 16 * 21)),
])
def test_coeff_count(native, h, w, mode, want):
    lib = native.load()
    n = C.c_uint64()
    native.check(lib.jds_coeff_count(h, w, native.SUBSAMPLING[mode], C.byref(n)))
    assert n.value == want


def test_odd_sizes_with_subsampling_are_supported(native):
    """Odd widths / heights floor the chroma plane like cv2.resize((W//2, H//2))
    (engines/color_space.py:44-49): 251x333 -> 125x166 chroma, padded to 128x168."""
    lib = native.load()
    n = C.c_uint64()
    assert lib.jds_coeff_count(251, 333, native.JDS_SUB_420, C.byref(n)) == native.JDS_OK
    assert n.value == 64 * (32 * 42 + 2 * 16 * 21)
    assert lib.jds_coeff_count(251, 333, native.JDS_SUB_422, C.byref(n)) == native.JDS_OK
    assert n.value == 64 * (32 * 42 + 2 * 32 * 21)
    ch, cw = C.c_int(), C.c_int()
    assert lib.jds_plane_dims(251, 333, native.JDS_SUB_420, C.byref(ch), C.byref(cw)) == native.JDS_OK
    assert (ch.value, cw.value) == (125, 166)
    # a subsampled plane cannot have width / height 0 (cv2.resize raises there too)
    assert lib.jds_coeff_count(8, 1, native.JDS_SUB_422, C.byref(n)) == native.JDS_ERR_INVALID


def test_params_validation_like_reference():
    from jpeg_dsp_studio_b200 import CompressionParams
    p = CompressionParams()
    assert (p.block_size, p.quality, p.subsampling_mode, p.use_prefilter) == (8, 50, '4:2:0', False)
    with pytest.raises(ValueError, match="Quality must be 1-100, got 0"):
        CompressionParams(quality=0)
    with pytest.raises(ValueError, match="Block size must be 4, 8, 16, or 32, got 7"):
        CompressionParams(block_size=7)
    CompressionParams(subsampling_mode="bogus")       # not validated at construction


def test_no_cpu_fallback():
    """Without a CUDA device the drop-in must raise, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jpeg_dsp_studio_b200 import CompressionParams, compress_reconstruct
    from jpeg_dsp_studio_b200._native import NativeError
    with pytest.raises(NativeError):
        compress_reconstruct(np.zeros((16, 16, 3), np.uint8), CompressionParams())


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "jpeg_dsp_studio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                if f == "__main__.py":
                    continue        # the CLI may use OpenCV for image FILE I/O only (optional import)
                assert not re.search(r"^\s*(from|import)\s+(scipy|cv2|skimage)\b", src, flags=re.M), f


def test_metric_finalisation_formulas():
    from jpeg_dsp_studio_b200.utils.metrics import psnr_from_sse
    assert psnr_from_sse(0, 100) == float("inf")
    a = np.random.default_rng(0).integers(0, 256, (40, 40, 3), dtype=np.uint8)
    b = np.random.default_rng(1).integers(0, 256, (40, 40, 3), dtype=np.uint8)
    from oracle.skimage_standin import peak_signal_noise_ratio
    sse = int(np.sum((a.astype(np.int64) - b.astype(np.int64)) ** 2))
    assert psnr_from_sse(sse, a.size) == peak_signal_noise_ratio(a, b, data_range=255)


def test_bitrate_finalisation_matches_reference_float32_arithmetic(golden_cases):
    """utils/metrics.py:63-92 does its arithmetic in float32 under NumPy 2; the host
    finalisation reproduces bpp / compression_ratio bit for bit from the exact integer
    counts whenever the total stays below 2**24 bits, and within 2e-7 above."""
    from types import SimpleNamespace
    from jpeg_dsp_studio_b200.utils.metrics import bitrate_from_partials
    from tests import cases as CS
    n_equal = 0
    for name, g in golden_cases.items():
        c = CS.BY_NAME[name]
        if c.big:
            continue
        img = c.image()
        o = P.compress_reconstruct(img, c.quality, c.mode, c.prefilter, want_metrics=False,
                                   want_maps=False)
        h, w = img.shape[:2]
        blocks = (-(-h // 8)) * (-(-w // 8))
        m = SimpleNamespace(luma_blocks=blocks, coeff_bits=o["exact_bits"] - 2 * blocks,
                            nnz=o["nonzero_coeffs"], total_coeffs=o["total_coeffs"])
        b = bitrate_from_partials(m, h, w)
        assert b["exact_bits"] == o["exact_bits"]
        if o["exact_bits"] < 2 ** 24:
            assert b["bpp"] == g["bpp"], name
            assert b["compression_ratio"] == g["compression_ratio"], name
            n_equal += 1
        else:
            assert abs(b["bpp"] - g["bpp"]) <= 2e-7 * g["bpp"]
    assert n_equal >= 15


def test_histogram_from_value_counts_equals_numpy_histogram_of_the_coefficients():
    """SURVEY 8f #2 host logic: binning the per-value counts (what the device returns) through
    np.histogram equals np.histogram of the full coefficient array (what matplotlib's ax.hist
    computes, gui/widgets/mpl_canvas.py:96) - counts and edges, any bin count."""
    from jpeg_dsp_studio_b200.engine import histogram_from_values
    from tests import cases as CS
    for seed, q, mode in ((1, 10, "4:2:0"), (2, 50, "4:4:4"), (3, 95, "4:2:2")):
        img = CS.rand_rgb(seed, 64, 80)
        coeffs = P.compress_reconstruct(img, q, mode, False, want_maps=False)["all_quantized_coeffs"]
        vals, cnt = np.unique(coeffs, return_counts=True)
        vh = np.zeros(2048, dtype=np.int64)
        vh[vals.astype(np.int64) + 1024] = cnt
        for bins in (7, 50, 64):
            c, e = histogram_from_values(vh, bins)
            wc, we = np.histogram(coeffs.flatten(), bins=bins)
            assert np.array_equal(c, wc) and np.array_equal(e, we)
    # a single distinct value (flat frame): numpy widens the range by +-0.5
    vh = np.zeros(2048, dtype=np.int64)
    vh[1024] = 4096
    c, e = histogram_from_values(vh, 50)
    wc, we = np.histogram(np.zeros(4096, dtype=np.int16), bins=50)
    assert np.array_equal(c, wc) and np.array_equal(e, we)


def test_sweep_quality_validation_keeps_the_reference_message():
    """Engine.sweep / sweep_records validate every point without a dataclass per point, with the
    reference's message (models/compression_params.py:16-17)."""
    from jpeg_dsp_studio_b200.engine import _check_qualities
    _check_qualities(list(range(1, 101)))
    for bad in (0, 101, -5):
        with pytest.raises(ValueError, match=f"Quality must be 1-100, got {bad}"):
            _check_qualities([50, bad, 60])


def test_table_finalisation_equals_per_record_formulas():
    """distributed.scalars_from_table (vectorised, one dict per row) against the per-record
    formulas of utils.metrics, incl. the float32 bit arithmetic and the all-zero-coefficient case;
    band partial records add up field by field."""
    from types import SimpleNamespace
    from jpeg_dsp_studio_b200 import distributed as D
    from jpeg_dsp_studio_b200.utils import metrics as M
    rng = np.random.default_rng(5)
    h, w = 2160, 3840
    rows = []
    for k in range(12):
        nnz = 0 if k == 3 else int(rng.integers(1, 12_000_000))
        m = SimpleNamespace(sse_rgb=int(rng.integers(0, 1 << 40)), sse_y=float(rng.random() * 1e9),
                            ssim_sum=[float(rng.random() * 8e6) for _ in range(4)], ssim_count=(h - 6) * (w - 6),
                            coeff_bits=6 * nnz + int(rng.integers(0, 5 * nnz + 1)), nnz=nnz,
                            total_coeffs=12441600, luma_blocks=129600)
        rows.append((m, D.record_from_metrics(k, k + 1, m)))
    table = D.scalars_from_table(np.stack([r for _, r in rows]), h, w)
    for (m, _), t in zip(rows, table):
        a, b = M.metrics_from_partials(m, h, w), M.bitrate_from_partials(m, h, w)
        assert t["psnr_rgb"] == a["psnr_rgb"] and t["psnr_y"] == a["psnr_y"]
        assert t["ssim_rgb"] == pytest.approx(a["ssim_rgb"], rel=1e-15) and t["ssim_y"] == a["ssim_y"]
        for key in ("estimated_bits", "exact_bits", "bpp", "compression_ratio", "nonzero_count", "total_coeffs"):
            assert t[key] == b[key], key
        assert all(type(v) in (int, float) for v in t.values())
    merged = D.merge_band_records([r for _, r in rows[:5]])
    assert merged[0] == rows[0][1][0] and merged[1] == rows[0][1][1]
    assert np.array_equal(merged[2:], np.sum([r[2:] for _, r in rows[:5]], axis=0))
