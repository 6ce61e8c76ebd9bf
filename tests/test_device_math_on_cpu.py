"""The product's device functions (csrc/jds_stages.cuh - the bodies of the CUDA
kernels) executed on the CPU through tests/emul and compared with the oracle.
This catches arithmetic / indexing bugs in the kernel source without a GPU; the
GPU parity tests (test_gpu_parity.py) then check the same code as real kernels."""

import ctypes as C

import numpy as np
import pytest

from oracle import numpy_port as P
from oracle import skimage_standin as SK
from tests import cases as CS
from tests.emul import build as emul_build

SUB = {"4:4:4": 0, "4:2:2": 1, "4:2:0": 2}


@pytest.fixture(scope="module")
def emul():
    return C.CDLL(emul_build.build())


def run_emul(lib, img, q, mode, pf, exact=1):
    H, W = img.shape[:2]
    sub = SUB[mode]
    hc = H // 2 if sub == 2 else H
    wc = W if sub == 0 else W // 2
    nb = (-(-H // 8)) * (-(-W // 8)) + 2 * (-(-hc // 8)) * (-(-wc // 8))
    recon = np.zeros((H, W, 3), np.uint8)
    coeffs = np.zeros(nb * 64, np.int16)
    ey, ergb = np.zeros((H, W)), np.zeros((H, W))
    bits, nnz, sse, ssey = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_double()
    hist = np.zeros(50, np.int64)
    img = np.ascontiguousarray(img)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emul_roundtrip(exact, H, W, q, sub, int(pf), vp(img), vp(recon), vp(coeffs), vp(ey),
                            vp(ergb), C.byref(bits), C.byref(nnz), vp(hist), C.byref(sse),
                            C.byref(ssey))
    assert rc == 0
    return dict(recon=recon, coeffs=coeffs, ey=ey, ergb=ergb, bits=bits.value, nnz=nnz.value,
                hist=hist, sse=sse.value, ssey=ssey.value)


SMALL = [c.name for c in CS.CASES if not c.big and max(c.image().shape) <= 512]


@pytest.mark.parametrize("name", SMALL)
def test_exact_device_math_bit_equal_to_oracle(emul, name):
    c = CS.BY_NAME[name]
    img = c.image()
    o = P.compress_reconstruct(img, c.quality, c.mode, c.prefilter, want_metrics=False)
    e = run_emul(emul, img, c.quality, c.mode, c.prefilter, exact=1)
    assert np.array_equal(e["coeffs"], o["all_quantized_coeffs"])
    assert np.array_equal(e["recon"], o["reconstructed_image"])
    assert np.array_equal(e["ey"], o["error_map_y"])
    assert np.array_equal(e["ergb"], o["error_map_rgb"])
    assert np.array_equal(e["hist"], o["quantized_histogram"])
    nblk = (-(-img.shape[0] // 8)) * (-(-img.shape[1] // 8))
    assert e["bits"] + 2 * nblk == o["exact_bits"]
    assert e["nnz"] == o["nonzero_coeffs"]
    d = img.astype(np.int64) - o["reconstructed_image"].astype(np.int64)
    assert e["sse"] == int(np.sum(d * d))


@pytest.mark.parametrize("name", ["rand250x334_q50_420_pf", "rand250x334_q50_422",
                                  "rand250x334_q90_444", "photo512_q75_420_pf"])
def test_fast_device_math_mismatch_rates(emul, name):
    """fp32 mode on natural / random content: mismatch rates of SURVEY §0.4."""
    c = CS.BY_NAME[name]
    img = c.image()
    o = P.compress_reconstruct(img, c.quality, c.mode, c.prefilter, want_metrics=False)
    e = run_emul(emul, img, c.quality, c.mode, c.prefilter, exact=0)
    assert np.mean(e["coeffs"] != o["all_quantized_coeffs"]) <= 1e-5
    assert np.mean(e["recon"] != o["reconstructed_image"]) <= 2e-4
    assert np.max(np.abs(e["recon"].astype(int) - o["reconstructed_image"].astype(int))) <= 1


def test_ssim_window_formula(emul):
    emul.emul_ssim_window.restype = C.c_double
    rng = np.random.default_rng(5)
    for trial in range(50):
        x = rng.integers(0, 256, (7, 7)).astype(np.float64)
        y = np.clip(x + rng.normal(0, 1 + trial, (7, 7)), 0, 255)
        if trial % 2:
            y = np.floor(y)
        want = SK.structural_similarity(x, y, data_range=255)     # 7x7 image -> one window
        vp = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
        for use_float, shift, tol in ((0, 0.0, 1e-12), (0, 128.0, 1e-12), (1, 128.0, 2e-5)):
            got = emul.emul_ssim_window(use_float, vp(x), vp(y), C.c_double(shift))
            assert abs(got - want) <= tol, (trial, use_float, shift, got, want)


def test_markstein_division_equals_ieee_division(emul):
    """exact mode quantises with RN(q0 + (X - q0 Q) RN(1/Q)) instead of the fp64 division
    subroutine; it must be the IEEE quotient for every table entry 1..255."""
    emul.emul_division_selftest.restype = C.c_long
    assert emul.emul_division_selftest(C.c_long(40000)) == 0


def test_strip_ssim_formula_keeps_precision_far_from_mid_grey(emul):
    """The per-window formula of k_ssim_strip (csrc/jds_ssim_formula.cuh, the same source the
    kernel compiles) on exact centred window sums: against the fp64 formula the error stays
    below 3e-7 per window on every content class - including flat content near black / white,
    where the naive fp32 form was off by up to 3e-5 per window with a systematic sign."""
    rng = np.random.default_rng(0)
    n = 20000

    def near(x, d):
        return np.clip(x + rng.integers(-d, d + 1, x.shape), 0, 255)

    classes = {
        "random": (rng.integers(0, 256, (n, 49)), rng.integers(0, 256, (n, 49))),
        "close": (lambda x: (x, near(x, 3)))(rng.integers(0, 256, (n, 49))),
        "dark": (lambda x: (x, near(x, 1)))(rng.integers(0, 6, (n, 49))),
        "bright": (lambda x: (x, near(x, 1)))(rng.integers(250, 256, (n, 49))),
        "flat levels": (lambda b: (near(b, 2), near(b, 2)))(np.repeat(rng.integers(2, 254, (n, 1)), 49, axis=1)),
    }
    f32 = np.float32
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for name, (x, y) in classes.items():
        x = x.astype(np.float64) - 128.0
        y = y.astype(np.float64) - 128.0
        sx, sy = x.sum(1), y.sum(1)
        sq, sc = (x * x + y * y).sum(1), (x * y).sum(1)
        args = [np.ascontiguousarray(v.astype(f32)) for v in (sx, sy, sq, sc)]
        assert all(np.array_equal(a.astype(np.float64), v) for a, v in zip(args, (sx, sy, sq, sc)))  # exact in fp32
        out = np.empty(n, dtype=f32)
        emul.emul_ssim_strip_formula(n, *[vp(a) for a in args], vp(out))
        ux, uy = sx / 49 + 128, sy / 49 + 128
        vxy = (sc / 49 - (sx / 49) * (sy / 49)) * 49 / 48
        vsum = (sq / 49 - (sx / 49) ** 2 - (sy / 49) ** 2) * 49 / 48
        ref = (2 * ux * uy + 6.5025) * (2 * vxy + 58.5225) / ((ux * ux + uy * uy + 6.5025) * (vsum + 58.5225))
        err = out.astype(np.float64) - ref
        assert np.abs(err).max() < 3e-7, (name, np.abs(err).max())
        assert abs(err.mean()) < 3e-8, (name, err.mean())
