"""The reference's own pipeline / subsampling tests, restated against the drop-in
(reference: tests/test_pipeline.py:9-49, tests/test_subsampling.py:10-70).  Same
inputs shapes, same assertions; seeded where the reference used the global RNG."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from jpeg_dsp_studio_b200 import CompressionParams, compress_reconstruct
    from jpeg_dsp_studio_b200.utils.test_images import (generate_colored_checkerboard,
                                                        generate_thin_stripes)
    return CompressionParams, compress_reconstruct, generate_colored_checkerboard, generate_thin_stripes


def _img(seed):
    return np.random.default_rng(seed).integers(0, 256, (64, 64, 3), dtype=np.uint8)


def test_quality_psnr_monotonic(api):
    CompressionParams, compress_reconstruct, _, _ = api
    image = _img(1)
    psnr_values = []
    for q in [10, 30, 50, 70, 90]:
        params = CompressionParams(quality=q, block_size=8, subsampling_mode='4:4:4')
        result, _ = compress_reconstruct(image, params)
        psnr_values.append(result.psnr_y)
    for i in range(len(psnr_values) - 1):
        assert psnr_values[i] <= psnr_values[i + 1] + 0.1


def test_perfect_reconstruction_high_quality(api):
    CompressionParams, compress_reconstruct, _, _ = api
    params = CompressionParams(quality=100, block_size=8, subsampling_mode='4:4:4')
    result, _ = compress_reconstruct(_img(2), params)
    assert result.psnr_y > 45.0


def test_subsampling_affects_quality(api):
    CompressionParams, compress_reconstruct, _, _ = api
    image = _img(3)
    r444, _ = compress_reconstruct(image, CompressionParams(quality=50, subsampling_mode='4:4:4'))
    r420, _ = compress_reconstruct(image, CompressionParams(quality=50, subsampling_mode='4:2:0'))
    assert r444.psnr_y >= r420.psnr_y


def test_compression_ratio_increases_with_lower_quality(api):
    CompressionParams, compress_reconstruct, _, _ = api
    image = _img(4)
    hi, _ = compress_reconstruct(image, CompressionParams(quality=90))
    lo, _ = compress_reconstruct(image, CompressionParams(quality=10))
    assert lo.compression_ratio >= hi.compression_ratio


def test_prefilter_reduces_aliasing(api):
    CompressionParams, compress_reconstruct, checker, _ = api
    board = checker(256)
    no_pf, _ = compress_reconstruct(board, CompressionParams(quality=50, block_size=8, subsampling_mode='4:2:0', use_prefilter=False))
    pf, _ = compress_reconstruct(board, CompressionParams(quality=50, block_size=8, subsampling_mode='4:2:0', use_prefilter=True))
    assert pf.ssim_rgb >= no_pf.ssim_rgb * 0.95


def test_subsampling_modes(api):
    CompressionParams, compress_reconstruct, _, _ = api
    image = _img(5)
    for mode in ['4:4:4', '4:2:2', '4:2:0']:
        result, _ = compress_reconstruct(image, CompressionParams(quality=50, block_size=8, subsampling_mode=mode))
        assert result.reconstructed_image.shape == image.shape


def test_thin_stripes_aliasing(api):
    CompressionParams, compress_reconstruct, _, stripes_gen = api
    stripes = stripes_gen(256, stripe_width=2)
    a, _ = compress_reconstruct(stripes, CompressionParams(quality=50, block_size=8, subsampling_mode='4:2:2', use_prefilter=False))
    b, _ = compress_reconstruct(stripes, CompressionParams(quality=50, block_size=8, subsampling_mode='4:2:2', use_prefilter=True))
    assert a.psnr_y > 0 and b.psnr_y > 0


def test_error_behaviour_matches_reference(api):
    """SURVEY §8a 'Error behaviour to preserve'."""
    CompressionParams, compress_reconstruct, _, _ = api
    img = _img(6)
    with pytest.raises(ValueError):
        CompressionParams(quality=0)
    with pytest.raises(ValueError):
        CompressionParams(quality=101)
    with pytest.raises(ValueError):
        CompressionParams(block_size=7)
    with pytest.raises(ValueError):                       # block_size 16 validates, then fails inside
        compress_reconstruct(img, CompressionParams(block_size=16))
    with pytest.raises(ValueError, match="Unknown subsampling mode"):
        compress_reconstruct(img, CompressionParams(subsampling_mode='4:1:1'))
    with pytest.raises(IndexError):
        compress_reconstruct(img[:, :, 0], CompressionParams())
    with pytest.raises(ValueError):
        compress_reconstruct(np.zeros((64, 64, 4), np.uint8), CompressionParams())
    with pytest.raises(ValueError):                       # smaller than the 7x7 SSIM window
        compress_reconstruct(np.zeros((6, 64, 3), np.uint8), CompressionParams(subsampling_mode='4:4:4'))
    # out-of-range selected block: fields stay None, index echoed (pipeline.py:137-138)
    _, inter = compress_reconstruct(img, CompressionParams(), (99, 99))
    assert inter.selected_block_idx == (99, 99) and inter.selected_block_dct is None
    _, inter = compress_reconstruct(img, CompressionParams(), (0, 9))   # aliases to block (1, 1)
    _, want = compress_reconstruct(img, CompressionParams(), (1, 1))
    assert np.array_equal(inter.selected_block_dct, want.selected_block_dct)
