"""Pin the oracle and the repo's input generators to the live reference.

Only runs where /root/reference exists (the build container); on the GPU box the
committed golden vectors (test_oracle_golden.py) carry the same evidence."""

import numpy as np
import pytest

from oracle import numpy_port as P
from oracle import reference_shim
from jpeg_dsp_studio_b200.utils import test_images as TI
from jpeg_dsp_studio_b200.utils import constants as K

pytestmark = pytest.mark.skipif(not reference_shim.available(),
                                reason="reference tree not present")


@pytest.fixture(scope="module")
def R():
    return reference_shim.load()


@pytest.mark.parametrize("name,args", [
    ("generate_colored_checkerboard", (512,)), ("generate_colored_checkerboard", (100,)),
    ("generate_thin_stripes", (256, 2)), ("generate_thin_stripes", (130, 3)),
    ("generate_gradient", (128,)), ("generate_text_edges", (200,)),
    ("generate_chroma_stripes", (250,)), ("generate_photo", (256,)),
])
def test_generators_match_reference(R, name, args):
    assert np.array_equal(getattr(TI, name)(*args), getattr(R.test_images, name)(*args))


def test_constants_match_reference(R):
    assert np.array_equal(K.JPEG_LUMA_Q50, R.constants.JPEG_LUMA_Q50)
    assert np.array_equal(K.ZIGZAG_ORDER, R.constants.ZIGZAG_ORDER)
    for q in range(1, 101):
        assert np.array_equal(P.scale_quant_matrix(q),
                              R.quantizer.scale_quant_matrix(R.constants.JPEG_LUMA_Q50, q))


def test_dct_restatement_bitwise(R):
    """A5/A6 against scipy (through the reference's dct2/idct2), bit for bit."""
    rng = np.random.default_rng(0)
    for _ in range(200):
        b = rng.uniform(-128, 128, (8, 8))
        assert np.array_equal(P.dct2_blocks(b), R.dct_engine.dct2(b))
        assert np.array_equal(P.idct2_blocks(b * 9), R.dct_engine.idct2(b * 9))


@pytest.mark.parametrize("shape,q,mode,pf,sel", [
    ((96, 120), 50, "4:2:0", True, (1, 2)),
    ((57 * 2, 33 * 2), 33, "4:2:2", True, (0, 0)),
    ((40, 56), 77, "4:4:4", False, (4, 6)),
    ((70, 70), 12, "4:2:0", False, (9, 0)),
])
def test_round_trip_matches_live_reference(R, shape, q, mode, pf, sel):
    img = np.random.default_rng(42).integers(0, 256, shape + (3,), dtype=np.uint8)
    res, inter = R.compress_reconstruct(
        img, R.CompressionParams(quality=q, subsampling_mode=mode, use_prefilter=pf), sel)
    o = P.compress_reconstruct(img, q, mode, pf, sel)
    assert np.array_equal(o["all_quantized_coeffs"], inter.all_quantized_coeffs)
    assert np.array_equal(o["reconstructed_image"], res.reconstructed_image)
    assert np.array_equal(o["error_map_y"], inter.error_map_y)
    assert np.array_equal(o["error_map_rgb"], inter.error_map_rgb)
    assert np.array_equal(o["quantized_histogram"], inter.quantized_histogram)
    assert (o["psnr_y"], o["ssim_y"], o["psnr_rgb"], o["ssim_rgb"]) == \
        (res.psnr_y, res.ssim_y, res.psnr_rgb, res.ssim_rgb)
    assert (o["bpp"], o["compression_ratio"]) == (res.bpp, res.compression_ratio)


# ---- stage functions one by one (the stand-alone operators of tests/test_stage_ops_gpu.py are
#      checked against these oracle functions on the GPU box) -----------------------------------
STAGE_SHAPES = [(16, 24), (33, 47), (48, 33), (250, 334), (7, 9)]


@pytest.mark.parametrize("shape", STAGE_SHAPES)
def test_stage_functions_match_reference(R, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    h, w = shape
    rgb = rng.uniform(0, 255, (h, w, 3))                       # arbitrary fp64 values, not only u8
    ref = R.color_space.rgb_to_ycbcr(rgb)
    Y, Cb, Cr = P.rgb_to_ycbcr(rgb)
    assert np.array_equal(np.stack([Y, Cb, Cr], -1), ref)
    ycc = rng.uniform(-20, 280, (h, w, 3))
    assert np.array_equal(P.ycbcr_to_rgb(ycc[..., 0], ycc[..., 1], ycc[..., 2]),
                          R.color_space.ycbcr_to_rgb(ycc))
    for mode in ("4:2:2", "4:2:0"):
        for pf in (False, True):
            rb, rr = R.color_space.subsample_chroma(ref[..., 1], ref[..., 2], mode, pf)
            ob = P.decimate_area(P.gaussian_blur_3x3(Cb) if pf else Cb, mode)
            orr = P.decimate_area(P.gaussian_blur_3x3(Cr) if pf else Cr, mode)
            assert np.array_equal(ob, rb) and np.array_equal(orr, rr)
            ub, ur = R.color_space.upsample_chroma(rb, rr, (h, w), 'bilinear')
            assert np.array_equal(P.upsample_linear(ob, h, w), ub)
            assert np.array_equal(P.upsample_linear(orr, h, w), ur)
    a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    b = np.clip(a.astype(np.int16) + rng.integers(-9, 10, (h, w, 3)), 0, 255).astype(np.uint8)
    assert P.psnr_ssim(a, b) == R.metrics.compute_psnr_ssim(a, b)
    q = rng.integers(-300, 300, 64 * 12).astype(np.int16)
    q[rng.random(q.size) < 0.6] = 0
    ref_b = R.metrics.estimate_bitrate_no_entropy(q, (16, 24), 8)
    ref_b.pop('label')
    assert P.bitrate_reference_arithmetic(q, (16, 24)) == ref_b


def test_block_processor_mirror_matches_reference(R):
    from jpeg_dsp_studio_b200.engines import block_processor as B
    rng = np.random.default_rng(3)
    for shape in ((16, 24), (13, 21), (8, 8), (9, 1 + 8)):
        ch = rng.uniform(0, 255, shape)
        pad, hw = B.pad_to_multiple(ch, 8)
        rpad, rhw = R.block_processor.pad_to_multiple(ch, 8)
        assert hw == rhw and np.array_equal(pad, rpad)
        mine, theirs = B.split_into_blocks(ch, 8), R.block_processor.split_into_blocks(ch, 8)
        assert len(mine) == len(theirs)
        for (i, j, blk), (ri, rj, rblk) in zip(mine, theirs):
            assert (i, j) == (ri, rj) and np.array_equal(blk, rblk)
        assert np.array_equal(B.merge_blocks(mine, shape, 8), R.block_processor.merge_blocks(theirs, shape, 8))
