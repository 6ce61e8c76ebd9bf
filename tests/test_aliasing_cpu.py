"""Chroma-aliasing demo without a GPU: the oracle restatement
(oracle/aliasing_port.py) against the golden vectors made by the unmodified reference worker,
against the live reference and live OpenCV where present, and the product's device arithmetic
(csrc/jds_alias.cuh compiled for the host, tests/emul) against the oracle."""

import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import aliasing_port as A
from oracle import reference_shim
from tests import cases as CS
from tests.conftest import parse_float

GOLD = os.path.join(os.path.dirname(__file__), "golden", "aliasing.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLD) as f:
        return {c["name"]: c for c in json.load(f)["cases"]}


def test_generators_match_oracle():
    from jpeg_dsp_studio_b200.engines import aliasing_demo as AD
    for name in ("generate_equiluminance_stripes", "generate_chroma_checkerboard", "generate_1px_checkerboard"):
        for size in (256, 64, 37):
            assert np.array_equal(getattr(AD, name)(size), getattr(A, name)(size)), (name, size)


@pytest.mark.parametrize("case", CS.ALIASING_CASES, ids=lambda c: c.name)
def test_oracle_matches_reference_golden(case, golden):
    g = golden[case.name]
    img = case.image()
    assert CS.sha(img) == g["input_sha256"]
    out = A.run_demo(img, case.quality)
    for k in ("recon_no_pf", "recon_pf", "diff_no_pf", "diff_pf"):
        assert CS.sha(out[k]) == g[k + "_sha256"], k
    for arm in ("metrics_no_pf", "metrics_pf"):
        for k, v in g[arm].items():
            assert out[arm][k] == pytest.approx(parse_float(v), abs=1e-12), (arm, k)


@pytest.mark.skipif(not reference_shim.available(), reason="reference tree not present")
def test_oracle_matches_live_reference_worker():
    R = reference_shim.load_aliasing_demo()
    for name in ("generate_equiluminance_stripes", "generate_chroma_checkerboard", "generate_1px_checkerboard"):
        assert np.array_equal(getattr(A, name)(48), getattr(R, name)(48))
    rng = np.random.default_rng(77)
    for shape, q in (((48, 64), 50), ((35, 41), 20), ((24, 17), 85)):
        img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        ref, mine = R.run_worker(img, q), A.run_demo(img, q)
        for k in ("recon_no_pf", "recon_pf", "diff_no_pf", "diff_pf"):
            assert np.array_equal(ref[k], mine[k]), (shape, k)
        for arm in ("metrics_no_pf", "metrics_pf"):
            for k in ref[arm]:
                assert mine[arm][k] == pytest.approx(ref[arm][k], abs=1e-12)
        m = R.compute_metrics(img, ref["recon_pf"])
        for k, v in A.compute_metrics(img, ref["recon_pf"]).items():
            assert v == pytest.approx(m[k], abs=1e-12)


def test_float32_kernels_match_live_opencv():
    """F1-F5 one by one against the OpenCV build of this image, on shapes that exercise the
    vector bodies and every tail (W mod 8, odd W)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    f32 = np.float32
    for H, W in ((16, 16), (21, 33), (37, 67), (40, 70), (9, 8), (12, 15), (64, 130), (50, 21)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ycc = cv2.cvtColor(img.astype(f32), cv2.COLOR_RGB2YCrCb)
        Y, Cr, Cb = A.rgb_to_ycrcb_f32(img.astype(f32))
        assert np.array_equal(Y, ycc[..., 0]) and np.array_equal(Cr, ycc[..., 1]) and np.array_equal(Cb, ycc[..., 2])
        assert np.array_equal(A.luma_u8(img), cv2.cvtColor(img, cv2.COLOR_RGB2YCrCb)[..., 0])
        plane = rng.uniform(-20, 280, (H, W)).astype(f32)
        assert np.array_equal(A.gaussian_blur_5x5_f32(plane), cv2.GaussianBlur(plane, (5, 5), 0.8))
        small = plane[::2, ::2]
        assert np.array_equal(A.resize_linear_f32(small, H, W),
                              cv2.resize(small, (W, H), interpolation=cv2.INTER_LINEAR))
        trip = np.stack([plane, rng.uniform(-20, 280, (H, W)).astype(f32),
                         rng.uniform(-20, 280, (H, W)).astype(f32)], axis=-1)
        assert np.array_equal(A.ycrcb_to_rgb_f32(trip[..., 0], trip[..., 1], trip[..., 2]),
                              cv2.cvtColor(trip, cv2.COLOR_YCrCb2RGB))
    k = cv2.getGaussianKernel(5, 0.8, cv2.CV_32F).ravel()
    assert (f32(k[0]), f32(k[1]), f32(k[2])) == (A.K0, A.K1, A.K2)


def test_fma32_is_correctly_rounded():
    """fma32 against exact rational arithmetic on values built to land on float32 ties."""
    from fractions import Fraction
    rng = np.random.default_rng(1)
    a = rng.uniform(-300, 300, 4000).astype(np.float32)
    b = rng.uniform(-2, 2, 4000).astype(np.float32)
    c = rng.uniform(-300, 300, 4000).astype(np.float32)
    # force near-ties: c = float32(x) - a*b with x halfway between two float32 values
    x = rng.uniform(1, 200, 2000).astype(np.float32)
    half = (x.astype(np.float64) + np.nextafter(x, np.float32(np.inf)).astype(np.float64)) / 2
    c[:2000] = (half - a[:2000].astype(np.float64) * b[:2000].astype(np.float64)).astype(np.float32)
    got = A.fma32(a, b, c)
    for i in range(0, 4000, 7):
        exact = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
        lo = np.float32(float(exact))            # float(Fraction) rounds correctly to fp64 ...
        # ... so decide between the two float32 neighbours of that value exactly
        cands = {np.float32(lo), np.nextafter(np.float32(lo), np.float32(np.inf)),
                 np.nextafter(np.float32(lo), np.float32(-np.inf))}
        best = min(cands, key=lambda v: (abs(Fraction(float(v)) - exact), int(np.float32(v).view(np.uint32)) & 1))
        assert got[i] == best, (i, a[i], b[i], c[i])


@pytest.fixture(scope="module")
def emul():
    from tests.emul import build
    lib = C.CDLL(build.build())
    lib.emul_alias_subsample.restype = C.c_int
    return lib


@pytest.mark.parametrize("shape", [(16, 16), (33, 70), (57, 75), (40, 9), (8, 8), (9, 15), (64, 128), (21, 33)])
def test_device_arithmetic_on_cpu_matches_oracle(emul, shape):
    H, W = shape
    img = np.random.default_rng(H * 1000 + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for pf in (0, 1):
        out = np.empty_like(img)
        assert emul.emul_alias_subsample(H, W, pf, vp(img), vp(out)) == 0
        assert np.array_equal(out, A.explicit_subsample_rgb(img, bool(pf))), (shape, pf)
    other = np.random.default_rng(3).integers(0, 256, img.shape, dtype=np.uint8)
    luma = np.empty((H, W), dtype=np.uint8)
    diff = np.empty_like(img)
    emul.emul_alias_luma_diff(C.c_size_t(H * W), vp(img), vp(other), vp(luma), vp(diff))
    assert np.array_equal(luma, A.luma_u8(img))
    assert np.array_equal(diff, A.compute_difference(img, other))


def test_worker_validates_like_the_reference():
    """quality outside 1..100 raises in CompressionParams before any GPU work (:152-157)."""
    from jpeg_dsp_studio_b200.engines.aliasing_demo import AliasingDemoWorker
    img = np.zeros((16, 16, 3), dtype=np.uint8)
    with pytest.raises(ValueError):
        AliasingDemoWorker(img, 0).run()
    with pytest.raises(ValueError):
        AliasingDemoWorker(img, 101).run()
