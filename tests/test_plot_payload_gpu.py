"""SURVEY 8f #2 - GUI plot payload: the histogram matplotlib would draw from
``all_quantized_coeffs`` (gui/compression_tab.py:662-667 -> gui/widgets/mpl_canvas.py:81-100)
and the x10 clipped error map (mpl_canvas.py:116-118), reduced on the device, against the
same reductions of the ORACLE's arrays."""
import numpy as np
import pytest

from tests import cases as CS

pytestmark = pytest.mark.gpu

SSIM_TOL = 1e-5


@pytest.fixture(scope="module")
def J():
    import jpeg_dsp_studio_b200 as J
    return J


@pytest.fixture(scope="module")
def oracle():
    from oracle import numpy_port
    return numpy_port


def _heat(err):
    return np.clip(err * 10, 0, 255).astype(np.uint8)          # mpl_canvas.py:118, truncated


PAYLOAD_CASES = [
    ("photo512_q50_420", lambda: CS.photo_tiled(512, 512), 50, "4:2:0", False),
    ("checker256_q10_420", lambda: CS.BY_NAME["cfg1_checker512_q10_420"].make()[:256, :256].copy(), 10, "4:2:0", False),
    ("rand250x334_q75_444", lambda: CS.rand_rgb(0, 250, 334), 75, "4:4:4", False),
    ("rand96x128_q30_422_pf", lambda: CS.rand_rgb(9, 96, 128), 30, "4:2:2", True),
    ("rand1080p_q75_420_pf", lambda: CS.rand_rgb(3, 1080, 1920), 75, "4:2:0", True),
    ("flat64_q50_420", lambda: np.full((64, 64, 3), 77, dtype=np.uint8), 50, "4:2:0", False),
]


@pytest.mark.parametrize("name,make,q,mode,pf", PAYLOAD_CASES, ids=[c[0] for c in PAYLOAD_CASES])
def test_plot_payload_exact_matches_reference_reductions(J, oracle, name, make, q, mode, pf):
    img = make()
    ref = oracle.compress_reconstruct(img, q, mode, pf)
    res, pay = J.plot_payload(img, J.CompressionParams(quality=q, subsampling_mode=mode, use_prefilter=pf),
                              want_heat_rgb=True)
    coeffs = ref["all_quantized_coeffs"]
    want_counts, want_edges = np.histogram(coeffs.flatten(), bins=50)    # what ax.hist computes
    assert pay.hist_counts.dtype == np.int64
    assert np.array_equal(pay.hist_counts, want_counts)
    assert np.array_equal(pay.hist_edges, want_edges)
    # per-value counts
    vals, cnt = np.unique(coeffs, return_counts=True)
    full = np.zeros(2048, dtype=np.int64)
    full[vals.astype(np.int64) + 1024] = cnt
    assert np.array_equal(pay.value_hist, full)
    assert np.array_equal(pay.error_heat_y, _heat(ref["error_map_y"]))
    assert np.array_equal(pay.error_heat_rgb, _heat(ref["error_map_rgb"]))
    assert np.array_equal(pay.reconstructed_image, ref["reconstructed_image"])
    assert res.original_image is img
    assert res.psnr_rgb == ref["psnr_rgb"]                  # integer squared error: exact
    assert abs(res.psnr_y - ref["psnr_y"]) <= 1e-9          # fp64 sum, different order
    assert res.nonzero_coeffs == ref["nonzero_coeffs"] and res.total_coeffs == ref["total_coeffs"]
    assert res.bpp == ref["bpp"]
    assert abs(res.ssim_rgb - ref["ssim_rgb"]) <= SSIM_TOL


def test_plot_payload_histogram_with_other_bin_counts(J, oracle):
    img = CS.rand_rgb(5, 64, 96)
    ref = oracle.compress_reconstruct(img, 20, "4:2:0", False, want_maps=False)
    for bins in (10, 50, 101):
        _, pay = J.plot_payload(img, J.CompressionParams(quality=20), bins=bins)
        c, e = np.histogram(ref["all_quantized_coeffs"], bins=bins)
        assert np.array_equal(pay.hist_counts, c) and np.array_equal(pay.hist_edges, e)


def test_plot_payload_device_tensor_and_fast_mode(J, oracle):
    import torch
    img = CS.rand_rgb(12, 128, 192)
    ref = oracle.compress_reconstruct(img, 60, "4:2:0", False)
    eng = J.get_engine()
    pay = eng.plot_payload(torch.from_numpy(img).cuda(), 60, "4:2:0", False, precision="exact")
    assert pay.error_heat_y.is_cuda and pay.reconstructed_image.is_cuda
    assert np.array_equal(pay.error_heat_y.cpu().numpy(), _heat(ref["error_map_y"]))
    assert np.array_equal(pay.hist_counts, np.histogram(ref["all_quantized_coeffs"], 50)[0])
    # fast (fp32) mode: the heat map may differ by one step where err*10 sits on an integer
    fast = eng.plot_payload(img, 60, "4:2:0", False, precision="fast")
    d = np.abs(fast.error_heat_y.astype(np.int16) - _heat(ref["error_map_y"]).astype(np.int16))
    frac = float((d > 0).mean())
    print(f"fast-mode heat map: {frac:.2e} of pixels differ, max step {int(d.max())}")
    assert frac < 0.02
    assert int(fast.hist_counts.sum()) == int(pay.hist_counts.sum())


def test_plot_payload_bytes_vs_intermediates(J):
    """The point of the row: bytes that cross PCIe per 4K frame."""
    h, w = 2160, 3840
    full = h * w * 3 + (h * w * 3 // 2) * 2 + 2 * h * w * 8        # recon + int16 coeffs + 2 fp64 maps
    payload = h * w * 3 + h * w + 2048 * 8                          # recon + uint8 heat + value counts
    assert payload * 4 < full
