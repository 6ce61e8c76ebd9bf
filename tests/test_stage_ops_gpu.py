"""The reference's stage functions as stand-alone GPU operators (engines/color_space.py,
utils/metrics.py mirrors; csrc/jds_ops.cu) against the oracle's restatement of the same
functions - which tests/test_oracle_vs_reference.py pins to the reference itself.
Bit-exact for everything fp64 / integer; SSIM within 1e-5."""
import numpy as np
import pytest

from tests import cases as CS

pytestmark = pytest.mark.gpu

SHAPES = [(16, 24), (33, 47), (48, 33), (250, 334), (7, 9), (1080, 1920), (541, 961)]


@pytest.fixture(scope="module")
def E():
    import jpeg_dsp_studio_b200.engines as E
    return E


@pytest.fixture(scope="module")
def oracle():
    from oracle import numpy_port
    return numpy_port


@pytest.mark.parametrize("shape", SHAPES, ids=[f"{h}x{w}" for h, w in SHAPES])
def test_colour_conversion_bit_exact(E, oracle, shape):
    rng = np.random.default_rng(shape[0] + shape[1])
    rgb = rng.uniform(0, 255, shape + (3,))
    Y, Cb, Cr = oracle.rgb_to_ycbcr(rgb)
    got = E.rgb_to_ycbcr(rgb)
    assert got.dtype == np.float64 and np.array_equal(got, np.stack([Y, Cb, Cr], -1))
    ycc = rng.uniform(-20, 280, shape + (3,))
    assert np.array_equal(E.ycbcr_to_rgb(ycc), oracle.ycbcr_to_rgb(ycc[..., 0], ycc[..., 1], ycc[..., 2]))
    # uint8-valued input, as the pipeline feeds it (engines/pipeline.py:25-28)
    img = CS.rand_rgb(1, *shape).astype(np.float64)
    Y, Cb, Cr = oracle.rgb_to_ycbcr(img)
    assert np.array_equal(E.rgb_to_ycbcr(img), np.stack([Y, Cb, Cr], -1))
    with pytest.raises(IndexError):
        E.rgb_to_ycbcr(np.zeros((8, 8)))


@pytest.mark.parametrize("shape", SHAPES, ids=[f"{h}x{w}" for h, w in SHAPES])
@pytest.mark.parametrize("mode", ["4:2:2", "4:2:0"])
@pytest.mark.parametrize("pf", [False, True], ids=["nopf", "pf"])
def test_subsample_upsample_bit_exact(E, oracle, shape, mode, pf):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    h, w = shape
    cb, cr = rng.uniform(0, 255, shape), rng.uniform(0, 255, shape)
    sb, sr = E.subsample_chroma(cb, cr, mode, pf)
    ob = oracle.decimate_area(oracle.gaussian_blur_3x3(cb) if pf else cb, mode)
    orr = oracle.decimate_area(oracle.gaussian_blur_3x3(cr) if pf else cr, mode)
    assert sb.shape == ob.shape and np.array_equal(sb, ob) and np.array_equal(sr, orr)
    ub, ur = E.upsample_chroma(sb, sr, (h, w), 'bilinear')
    assert np.array_equal(ub, oracle.upsample_linear(ob, h, w))
    assert np.array_equal(ur, oracle.upsample_linear(orr, h, w))


def test_subsample_444_and_errors(E):
    cb = np.arange(64, dtype=np.float64).reshape(8, 8)
    a, b = E.subsample_chroma(cb, cb + 1, '4:4:4')
    assert np.array_equal(a, cb) and a is not cb and np.array_equal(b, cb + 1)
    with pytest.raises(ValueError, match="Unknown subsampling mode"):
        E.subsample_chroma(cb, cb, '4:1:1')
    with pytest.raises(NotImplementedError):
        E.upsample_chroma(cb, cb, (16, 16), 'nearest')


def test_general_enlargement_factors(E, oracle):
    rng = np.random.default_rng(8)
    for _ in range(12):
        h, w = int(rng.integers(2, 60)), int(rng.integers(2, 60))
        H, W = int(rng.integers(h, 4 * h)), int(rng.integers(w, 4 * w))
        p = rng.uniform(0, 255, (h, w))
        u, _ = E.upsample_chroma(p, p, (H, W))
        assert np.array_equal(u, oracle.upsample_linear(p, H, W)), (h, w, H, W)


@pytest.mark.parametrize("shape", [(64, 64), (33, 47), (250, 334), (1080, 1920), (7, 7)],
                         ids=lambda s: f"{s[0]}x{s[1]}")
def test_compute_psnr_ssim(oracle, shape):
    from jpeg_dsp_studio_b200.utils.metrics import compute_psnr_ssim
    rng = np.random.default_rng(shape[1])
    a = CS.photo_tiled(*shape) if shape[0] > 100 else CS.rand_rgb(2, *shape)
    b = np.clip(a.astype(np.int16) + rng.integers(-12, 13, a.shape), 0, 255).astype(np.uint8)
    got, want = compute_psnr_ssim(a, b), oracle.psnr_ssim(a, b)
    assert set(got) == {'psnr_rgb', 'ssim_rgb', 'psnr_y', 'ssim_y'}
    assert got['psnr_rgb'] == want['psnr_rgb']
    assert abs(got['psnr_y'] - want['psnr_y']) <= 1e-9
    assert abs(got['ssim_rgb'] - want['ssim_rgb']) <= 1e-5
    assert abs(got['ssim_y'] - want['ssim_y']) <= 1e-5
    same = compute_psnr_ssim(a, a)
    assert same['psnr_rgb'] == float('inf') and abs(same["ssim_y"] - 1.0) <= 1e-5
    with pytest.raises(ValueError):
        compute_psnr_ssim(a[:5], b[:5])


def test_compute_psnr_ssim_on_device_tensors(oracle):
    import torch
    from jpeg_dsp_studio_b200.utils.metrics import compute_psnr_ssim
    a = CS.rand_rgb(3, 96, 128)
    b = np.clip(a.astype(np.int16) + 3, 0, 255).astype(np.uint8)
    got = compute_psnr_ssim(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    want = oracle.psnr_ssim(a, b)
    assert got['psnr_rgb'] == want['psnr_rgb'] and abs(got['ssim_rgb'] - want['ssim_rgb']) <= 1e-5


def test_estimate_bitrate_no_entropy(oracle):
    from jpeg_dsp_studio_b200.utils.metrics import estimate_bitrate_no_entropy
    rng = np.random.default_rng(4)
    for shape, density in (((64, 64), 0.3), ((250, 334), 0.05), ((1080, 1920), 0.5), ((16, 16), 0.0)):
        nblk = (-(-shape[0] // 8)) * (-(-shape[1] // 8))
        q = rng.integers(-1016, 1017, nblk * 64 * 3).astype(np.int16)
        q[rng.random(q.size) >= density] = 0
        got = estimate_bitrate_no_entropy(q, shape, 8)
        want = oracle.bitrate_reference_arithmetic(q, shape)
        for k in ('estimated_bits', 'bpp', 'compression_ratio', 'nonzero_count', 'total_coeffs', 'label'):
            if k in want:
                assert got[k] == want[k], (shape, k, got[k], want[k])
    # int64 input: NumPy's log2 is float64 there, the sums are exact
    q64 = q.astype(np.int64)
    got = estimate_bitrate_no_entropy(q64, (16, 16), 8)
    nz = q64[q64 != 0]
    bits = 2 * 4 + 6 * nz.size + int(np.sum(np.ceil(np.log2(np.abs(nz) + 1)) + 1)) if nz.size else 2 * 4
    assert got['estimated_bits'] == bits


@pytest.mark.parametrize("lo,hi", [(0, 6), (250, 256), (0, 2), (126, 131), (0, 256)])
def test_ssim_accuracy_far_from_mid_grey(lo, hi):
    """fp32 window sums are centred at 128: on flat content near black / white the covariance
    and variance terms are small differences of ~4e7-sized products.  The compensated formula
    (jds_ssim.cu) keeps the integer channels exact (was 1.6e-5 on a dark frame) and Y within
    the 1e-5 tolerance."""
    from jpeg_dsp_studio_b200.utils.metrics import compute_psnr_ssim
    from oracle import numpy_port as P
    rng = np.random.default_rng(lo * 1000 + hi)
    a = rng.integers(lo, hi, (272, 480, 3), dtype=np.uint8)
    b = np.clip(a.astype(np.int16) + rng.integers(-1, 2, a.shape), 0, 255).astype(np.uint8)
    g, o = compute_psnr_ssim(a, b), P.psnr_ssim(a, b)
    assert abs(g["ssim_rgb"] - o["ssim_rgb"]) <= 1e-6
    assert abs(g["ssim_y"] - o["ssim_y"]) <= 1e-5
    assert g["psnr_rgb"] == o["psnr_rgb"]
    assert abs(g["psnr_y"] - o["psnr_y"]) <= 1e-9
