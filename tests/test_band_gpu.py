"""Tile-band sharding of one frame (jds_roundtrip_band): the bands of any
world size, run one after the other on this GPU, must reassemble to exactly what the whole-frame
round trip produces - pixels and coefficients bit for bit, integer partials exactly, the fp64
SSIM / Y sums up to summation order."""

import numpy as np
import pytest

from tests import cases as CS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def J():
    import jpeg_dsp_studio_b200 as J
    return J


def _blocks(h, w, mode):
    v = 2 if mode == "4:2:0" else 1
    hd = 1 if mode == "4:4:4" else 2
    nb = lambda n: (n + 7) // 8
    return nb(h) * nb(w), nb(h // v) * nb(w // hd)


def _check(J, img, q, mode, pf, precision, worlds):
    from jpeg_dsp_studio_b200 import distributed as D
    eng = J.get_engine()
    h, w = img.shape[:2]
    full = eng.roundtrip(img, q, mode, pf, precision=precision, want_coeffs=True)
    want_rec = D.record_from_metrics(0, q, full.metrics)
    ny, nc = _blocks(h, w, mode)
    fy, fcb, fcr = np.split(np.asarray(full.coeffs), [64 * ny, 64 * (ny + nc)])
    for world in worlds:
        bounds = D.band_bounds(h, world)
        assert [b for b in bounds if b][0][0] == 0 and [b for b in bounds if b][-1][1] == h
        outs = [eng.roundtrip_band(img, b[0], b[1], q, mode, pf, precision=precision, want_coeffs=True)
                for b in bounds if b]
        recon = np.concatenate([o.recon for o in outs], axis=0)
        assert np.array_equal(recon, full.recon), (world, mode, pf)
        parts = [[], [], []]
        for o, b in zip(outs, [b for b in bounds if b]):
            bny, bnc = int(o.metrics.luma_blocks), (int(o.metrics.total_coeffs) // 64 - int(o.metrics.luma_blocks)) // 2
            y, cb, cr = np.split(np.asarray(o.coeffs), [64 * bny, 64 * (bny + bnc)])
            for lst, a in zip(parts, (y, cb, cr)):
                lst.append(a)
        for lst, a in zip(parts, (fy, fcb, fcr)):
            assert np.array_equal(np.concatenate(lst), a), (world, mode, pf)
        got = D.merge_band_records([D.record_from_metrics(0, q, o.metrics) for o in outs])
        names = D.RECORD_FIELDS
        for k, name in enumerate(names):
            if name == "sse_y":
                # the whole-frame path takes the Y error from the SSIM kernel's fp32 window sums
                # (sum x^2 + y^2 - 2xy), the band path from the fp64 SSE kernel: 2e-5 relative on
                # low-error content = 1e-4 dB of PSNR_y (target: 1e-3 dB)
                assert got[k] == pytest.approx(want_rec[k], rel=1e-4), (name, world)
            elif name in ("ssim_r", "ssim_g", "ssim_b", "ssim_y"):
                # fp32 sliding-window sums: the rounding depends on where a strip segment starts
                assert got[k] == pytest.approx(want_rec[k], rel=2e-6, abs=1e-6), (name, world)
            else:
                assert got[k] == want_rec[k], (name, world)
        a, b = D.scalars_from_record(got, h, w), D.scalars_from_record(want_rec, h, w)
        for key in ("psnr_rgb", "bpp", "compression_ratio", "nonzero_count"):
            assert a[key] == b[key]
        assert a["ssim_rgb"] == pytest.approx(b["ssim_rgb"], abs=2e-6)       # see above; target 1e-5


@pytest.mark.parametrize("shape,q,mode,pf", [
    ((160, 224), 50, "4:2:0", False), ((160, 224), 50, "4:2:0", True), ((250, 334), 35, "4:2:2", True),
    ((129, 96), 80, "4:4:4", False), ((250, 335), 20, "4:2:2", False), ((96, 64), 10, "4:2:0", True),
    ((52, 80), 60, "4:2:0", False),
])
def test_bands_reassemble_exact_mode(J, shape, q, mode, pf):
    _check(J, CS.rand_rgb(shape[0] + q, *shape), q, mode, pf, "exact", (1, 2, 3, 5, 8))


def test_bands_reassemble_natural_content(J):
    _check(J, CS.photo_tiled(272, 320), 75, "4:2:0", True, "exact", (2, 4))


@pytest.mark.parametrize("pf", [False, True])
def test_bands_4k_fast_mode(J, pf):
    """BASELINE's 4K frame over 2 / 8 ranks' bands with the fused fp32 kernels: a band is a
    16-aligned frame of its own, so the same kernels serve band and whole frame"""
    _check(J, CS.photo_tiled(2160, 3840), 50, "4:2:0", pf, "fast", (2, 8))


def test_band_device_input_and_errors(J):
    import torch
    from jpeg_dsp_studio_b200 import distributed as D
    eng = J.get_engine()
    img = CS.rand_rgb(5, 192, 256)
    full = eng.roundtrip(img, 40, "4:2:0", False, precision="exact")
    d = torch.from_numpy(img).cuda()
    out = eng.roundtrip_band(d, 64, 128, 40, "4:2:0", False, precision="exact")
    assert out.recon.is_cuda and np.array_equal(out.recon.cpu().numpy(), full.recon[64:128])
    single = D.frame_banded(eng, img, 40, "4:2:0", False)          # no process group: one band
    assert np.array_equal(single["recon"], full.recon)
    assert single["scalars"]["psnr_rgb"] == full.scalars["psnr_rgb"]
    with pytest.raises(Exception, match="multiples of 16"):
        eng.roundtrip_band(img, 8, 64, 40, "4:2:0", False)
    with pytest.raises(Exception, match="multiples of 16"):
        eng.roundtrip_band(img, 0, 100, 40, "4:2:0", False)
    odd = CS.rand_rgb(6, 99, 64)
    with pytest.raises(Exception, match="odd height"):
        eng.roundtrip_band(odd, 0, 48, 40, "4:2:0", False)
    whole = eng.roundtrip_band(odd, 0, 99, 40, "4:2:0", False)     # the whole frame is always fine
    assert np.array_equal(whole.recon, eng.roundtrip(odd, 40, "4:2:0", False).recon)
