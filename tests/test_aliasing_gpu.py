"""Chroma-aliasing demo on the GPU, through the C ABI (jds_aliasing_demo /
jds_aliasing_metrics): every golden case of the unmodified reference worker, bit for bit."""

import json
import os

import numpy as np
import pytest

from oracle import aliasing_port as A
from tests import cases as CS
from tests.conftest import parse_float

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "aliasing.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLD) as f:
        return {c["name"]: c for c in json.load(f)["cases"]}


@pytest.fixture(scope="module")
def AD():
    from jpeg_dsp_studio_b200.engines import aliasing_demo
    return aliasing_demo


@pytest.mark.parametrize("case", CS.ALIASING_CASES, ids=lambda c: c.name)
def test_worker_matches_reference_golden(AD, case, golden):
    g = golden[case.name]
    img = case.image()
    msgs = []
    out = AD.AliasingDemoWorker(img, case.quality, progress=msgs.append).run()
    assert out["original"] is img
    assert msgs == ["Processing without prefilter (true decimation)...",
                    "Processing with Gaussian prefilter..."]
    for k in ("recon_no_pf", "recon_pf", "diff_no_pf", "diff_pf"):
        assert out[k].dtype == np.uint8 and out[k].shape == img.shape
        assert CS.sha(out[k]) == g[k + "_sha256"], k
    for arm in ("metrics_no_pf", "metrics_pf"):
        for k, v in g[arm].items():
            tol = 1e-5 if k.startswith("ssim") else 1e-9     # PSNR: integer squared errors
            assert out[arm][k] == pytest.approx(parse_float(v), abs=tol), (arm, k)


@pytest.mark.parametrize("shape", [(64, 80), (41, 53), (8, 8), (1080, 1920)])
def test_subsampled_frame_and_oracle(AD, shape):
    """the uint8 frame handed to the hot path (float32 OpenCV arithmetic) against the oracle,
    host and device input, both arms"""
    import torch
    import jpeg_dsp_studio_b200 as J
    img = CS.rand_rgb(sum(shape), *shape)
    eng = J.get_engine()
    for pf in (False, True):
        want = A.explicit_subsample_rgb(img, pf)
        got = eng.aliasing_demo_arm(img, 50, pf, want_subsampled=True)
        assert np.array_equal(got["subsampled"], want), (shape, pf)
        dgot = eng.aliasing_demo_arm(torch.from_numpy(img).cuda(), 50, pf, want_subsampled=True)
        assert dgot["recon"].is_cuda
        assert np.array_equal(dgot["subsampled"].cpu().numpy(), want)
        assert np.array_equal(dgot["recon"].cpu().numpy(), got["recon"])
        assert np.array_equal(dgot["diff"].cpu().numpy(), A.compute_difference(img, got["recon"]))
        if max(shape) <= 128:
            ref = A.process_with_explicit_subsample(img, 50, pf)
            assert np.array_equal(got["recon"], ref)


def test_compute_metrics_standalone(AD):
    a = CS.rand_rgb(90, 72, 96)
    b = np.clip(a.astype(np.int16) + np.random.default_rng(91).integers(-9, 10, a.shape), 0, 255).astype(np.uint8)
    got, want = AD.compute_metrics(a, b), A.compute_metrics(a, b)
    for k in want:
        tol = 1e-5 if k.startswith("ssim") else 1e-9
        assert got[k] == pytest.approx(want[k], abs=tol), k
    same = AD.compute_metrics(a, a)
    assert same["psnr_rgb"] == float("inf") and same["psnr_y"] == float("inf")
    assert same["ssim_y"] == pytest.approx(1.0, abs=1e-6)
    with pytest.raises(ValueError):
        AD.compute_metrics(a, b[:-8])


def test_fast_precision_is_reported_not_hidden(AD, record_property):
    """the hot-path stage may run in fp32: the front end stays bit-exact, the reconstruction
    differs from the exact arm in a bounded fraction of samples on natural content"""
    from jpeg_dsp_studio_b200.utils import test_images as TI
    img = TI.generate_photo(256)
    ex = AD.AliasingDemoWorker(img, 50).run()
    fa = AD.AliasingDemoWorker(img, 50, precision="fast").run()
    for k in ("recon_no_pf", "recon_pf"):
        mism = float(np.mean(ex[k] != fa[k]))
        record_property(f"fast_mismatch_{k}", mism)
        assert mism <= 2e-3
    for arm in ("metrics_no_pf", "metrics_pf"):
        assert abs(ex[arm]["psnr_y"] - fa[arm]["psnr_y"]) <= 1e-2
        assert abs(ex[arm]["ssim_y"] - fa[arm]["ssim_y"]) <= 1e-4


def test_rejects_tiny_frames(AD):
    from jpeg_dsp_studio_b200._native import NativeError
    with pytest.raises(NativeError):
        AD.AliasingDemoWorker(np.zeros((4, 4, 3), np.uint8), 50).run()
