"""The reference's DCT unit tests (tests/test_dct.py:8-47) restated against the drop-in
operators, plus bit-equality of the operators with the oracle."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    from jpeg_dsp_studio_b200 import engines
    return engines


def test_dct_idct_invertibility(E):
    block = np.random.default_rng(1).random((8, 8)) * 255
    recovered = E.idct2(E.dct2(block - 128.0)) + 128.0
    assert np.allclose(block, recovered, atol=1e-10)


def test_encode_decode_block_invertibility(E):
    block = np.random.default_rng(2).random((8, 8)) * 255
    assert np.allclose(block, E.decode_block(E.encode_block(block)), atol=1e-8)


def test_level_shift_reduces_dc(E):
    block = np.ones((8, 8)) * 200
    assert abs(E.dct2(block - 128.0)[0, 0]) < abs(E.dct2(block)[0, 0])


def test_energy_preservation(E):
    shifted = np.random.default_rng(3).random((8, 8)) * 255 - 128.0
    assert np.isclose(np.sum(shifted ** 2), np.sum(E.dct2(shifted) ** 2), rtol=1e-10)


def test_constant_block_dct(E):
    d = E.dct2(np.ones((8, 8)) * 128 - 128.0)
    assert np.allclose(d[0, 1:], 0, atol=1e-10) and np.allclose(d[1:, :], 0, atol=1e-10)


def test_block_operators_bit_equal_to_oracle(E):
    from oracle import numpy_port as P
    rng = np.random.default_rng(4)
    x = rng.uniform(0, 255, (500, 8, 8))
    assert np.array_equal(E.dct2(x - 128.0), P.dct2_blocks(x - 128.0))
    assert np.array_equal(E.encode_block(x), P.dct2_blocks(x - 128.0))
    c = P.dct2_blocks(x - 128.0)
    assert np.array_equal(E.idct2(c), P.idct2_blocks(c))
    assert np.array_equal(E.decode_block(c * 1.3), np.clip(P.idct2_blocks(c * 1.3) + 128.0, 0, 255))
    for q in (1, 10, 50, 93, 100):
        Q = E.scale_quant_matrix(E.JPEG_LUMA_Q50, q)
        assert np.array_equal(Q, P.scale_quant_matrix(q))
        qz = E.quantize(c, Q)
        assert qz.dtype == np.int16 and np.array_equal(qz, np.round(c / Q).astype(np.int16))
        assert np.array_equal(E.dequantize(qz, Q), qz.astype(np.float64) * Q)
    with pytest.raises(ValueError):
        E.quantize(np.zeros((16, 16)), E.scale_quant_matrix(E.JPEG_LUMA_Q50, 50))
