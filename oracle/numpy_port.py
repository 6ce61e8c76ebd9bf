"""Vectorised NumPy restatement of the reference round trip (TEST INFRASTRUCTURE).

Follows ``/root/reference/engines/pipeline.py:17-167`` stage by stage, and for the
third-party arithmetic the reference calls (SciPy/ducc0 DCT, OpenCV GaussianBlur /
INTER_AREA / IPP INTER_LINEAR) the operation order of SURVEY.md Appendix A.  NumPy
never contracts ``a*b+c`` into an FMA, so every ``+ - * /`` below is one IEEE-754
binary64 operation; the two places the reference's libraries DO use an FMA are
written with ``_fma`` (exact, via error-free transformations is overkill here:
``math.fma`` is used through ``np.frompyfunc`` only on the few FMA sites, or the
long-double path when available - see ``_fma``).

Parity status: pinned - see oracle/__init__.py and tests/test_oracle_vs_reference.py.
"""

import math

import numpy as np

from . import skimage_standin as _sk

# ---------------------------------------------------------------------------
# constants
# ---------------------------------------------------------------------------

#: /root/reference/utils/constants.py:6-15
JPEG_LUMA_Q50 = np.array([
    [16, 11, 10, 16, 24, 40, 51, 61],
    [12, 12, 14, 19, 26, 58, 60, 55],
    [14, 13, 16, 24, 40, 57, 69, 56],
    [14, 17, 22, 29, 51, 87, 80, 62],
    [18, 22, 37, 56, 68, 109, 103, 77],
    [24, 35, 55, 64, 81, 104, 113, 92],
    [49, 64, 78, 87, 103, 121, 120, 101],
    [72, 92, 95, 98, 112, 100, 103, 99]], dtype=np.float64)

_h = float.fromhex
#: ducc0 UnityRoots twiddles for N=8 - libm-derived, not
#: correctly-rounded cosines.
TW = (_h("0x1.f6297cff75cb0p-1"), _h("0x1.d906bcf328d46p-1"),
      _h("0x1.a9b66290ea1a3p-1"), _h("0x1.6a09e667f3bccp-1"),
      _h("0x1.1c73b39ae68c8p-1"), _h("0x1.87de2a6aea963p-2"),
      _h("0x1.8f8b83c69a60ap-3"))
WR = _h("0x1.6a09e667f3bccp-1")
WI = _h("0x1.6a09e667f3bcdp-1")
S2 = _h("0x1.6a09e667f3bcdp+0")
HALF_S2 = _h("0x1.6a09e667f3bcdp-1")
#: cv2.getGaussianKernel(3, 0.75) in fp64 (Appendix A2)
K_E = _h("0x1.ce0cac8ce5377p-3")
K_C = _h("0x1.18f9a9b98d643p-1")

MODES = ("4:4:4", "4:2:2", "4:2:0")


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------

_fma_ufunc = np.frompyfunc(math.fma, 3, 1) if hasattr(math, "fma") else None


def _fma(a, b, c):
    """Correctly-rounded fused multiply-add on fp64 arrays.

    Python 3.13 has ``math.fma``; on 3.12 fall back to an exact computation with
    Dekker/Knuth error-free transformations (TwoProduct via Veltkamp splitting,
    then a correctly rounded three-term sum is not needed: all FMA sites here have
    operands far from overflow/underflow, and the result of ``a*b + c`` is rounded
    once from the exact value ``p + e + c``).
    """
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    if _fma_ufunc is not None:
        return _fma_ufunc(a, b, c).astype(np.float64)
    # np.longdouble on x86-64 is the 80-bit x87 type (64-bit significand): the
    # product of two doubles needs 106 bits, so that is not enough on its own.
    # Exact path: p + e = a*b exactly (TwoProduct), then round(p + e + c) once.
    a, b, c = np.broadcast_arrays(a, b, c)
    split = 134217729.0  # 2**27 + 1
    p = a * b
    ta = split * a
    ah = ta - (ta - a)
    al = a - ah
    tb = split * b
    bh = tb - (tb - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl          # a*b == p + e exactly
    # exact sum of three doubles rounded once: use TwoSum(p, c) -> s + t, then the
    # result is round(s + (t + e)) when |t + e| is representable enough; handle the
    # rare double-rounding case with a sticky correction (Boldo-Melquiond).
    s = p + c
    bb = s - p
    t = (p - (s - bb)) + (c - bb)                               # p + c == s + t exactly
    # now need round(s + t + e); u = t + e may round:
    u = t + e
    bb2 = u - t
    v = (t - (u - bb2)) + (e - bb2)                             # t + e == u + v exactly
    # round-to-odd emulation of u (sticky from v) so the final add rounds correctly
    res = s + u
    need = v != 0.0
    if np.any(need):
        uu = u[need]
        vv = v[need]
        # if u is "even" in its last bit and v != 0, nudge u one ulp toward v
        bits = uu.view(np.int64)
        odd = (bits & 1).astype(bool)
        toward = np.where(vv > 0, np.inf, -np.inf)
        nudged = np.nextafter(uu, toward)
        # round-to-odd: keep uu if its last bit is odd, else take the odd neighbour
        ro = np.where(odd, uu, nudged)
        res = res.copy()
        res[need] = s[need] + ro
    return res


def scale_quant_matrix(quality):
    """/root/reference/engines/quantizer.py:7-19."""
    quality = int(np.clip(quality, 1, 100))
    scale = 5000.0 / quality if quality < 50 else 200.0 - 2.0 * quality
    Q = np.floor((JPEG_LUMA_Q50 * scale + 50.0) / 100.0)
    return np.clip(Q, 1, 255).astype(np.float64)


def rgb_to_ycbcr(rgb_f64):
    """A1 - /root/reference/engines/color_space.py:8-14 (NumPy left-to-right)."""
    R, G, B = rgb_f64[..., 0], rgb_f64[..., 1], rgb_f64[..., 2]
    Y = 0.299 * R + 0.587 * G + 0.114 * B
    Cb = -0.168736 * R - 0.331264 * G + 0.5 * B + 128.0
    Cr = 0.5 * R - 0.418688 * G - 0.081312 * B + 128.0
    return Y, Cb, Cr


def gaussian_blur_3x3(x):
    """A2 - cv2.GaussianBlur(f64, (3,3), sigmaX=0.75), BORDER_REFLECT_101."""
    H, W = x.shape
    if H < 2 or W < 2:
        raise NotImplementedError("prefilter restatement needs H, W >= 2")
    xp = np.empty((H, W + 2), dtype=np.float64)
    xp[:, 1:-1] = x
    xp[:, 0] = x[:, 1]
    xp[:, -1] = x[:, W - 2]
    left, mid, right = xp[:, :-2], xp[:, 1:-1], xp[:, 2:]
    W4 = 4 * (W // 4)
    r = np.empty((H, W), dtype=np.float64)
    if W4:
        r[:, :W4] = _fma(K_E, right[:, :W4],
                         _fma(K_C, mid[:, :W4], K_E * left[:, :W4]))
    if W4 < W:
        r[:, W4:] = ((K_E * left[:, W4:]) + (K_C * mid[:, W4:])) + (K_E * right[:, W4:])
    rp = np.empty((H + 2, W), dtype=np.float64)
    rp[1:-1] = r
    rp[0] = r[1]
    rp[-1] = r[H - 2]
    return (K_C * rp[1:-1]) + (K_E * (rp[2:] + rp[:-2]))


def area_tab(ssize, dsize):
    """OpenCV 4.13 imgproc/resize.cpp computeResizeAreaTab: the (dst, src, weight) taps of
    cv2.resize(INTER_AREA) at a non-integer shrink factor.  Coordinates are fp64, the weight
    is rounded to float32 (DecimateAlpha::alpha is a float).  [verified against cv2 4.13.0
    on 200 random odd/even plane sizes, 0 mismatches]"""
    import math
    scale = ssize / dsize
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area_general(x, dh, dw):
    """cv2.resize(f64 plane, (dw, dh), INTER_AREA) when a shrink factor is not an integer
    (OpenCV ResizeArea_Invoker<double, double>): per source row buf[dx] = buf[dx] + S*alpha
    over the x taps in order (buf starts at 0), then per destination row
    sum = beta*buf for its first y tap and sum = sum + beta*buf for the following ones.
    No FMA; alpha / beta are float32 values widened to fp64."""
    sh, sw = x.shape
    xt, yt = area_tab(sw, dw), area_tab(sh, dh)
    buf = np.zeros((sh, dw), dtype=np.float64)
    for dx, s, a in xt:
        buf[:, dx] = buf[:, dx] + x[:, s] * np.float64(a)
    out = np.zeros((dh, dw), dtype=np.float64)
    first = np.ones(dh, dtype=bool)
    for dy, s, b in yt:
        if first[dy]:
            out[dy] = np.float64(b) * buf[s]
            first[dy] = False
        else:
            out[dy] = out[dy] + np.float64(b) * buf[s]
    return out


def decimate_area(x, mode):
    """A3 - cv2.resize(INTER_AREA) to (W//2, H) or (W//2, H//2)
    (/root/reference/engines/color_space.py:44-49).  Even sizes take OpenCV's integer-factor
    path; an odd width (or height under 4:2:0) takes the general path for BOTH axes."""
    H, W = x.shape
    if W % 2 or (mode == "4:2:0" and H % 2):
        return resize_area_general(x, H // 2 if mode == "4:2:0" else H, W // 2)
    if mode == "4:2:2":
        return (x[:, 0::2] + x[:, 1::2]) * 0.5
    a, b = x[0::2, 0::2], x[0::2, 1::2]
    c, d = x[1::2, 0::2], x[1::2, 1::2]
    return (((a + b) + c) + d) * 0.25


def reflect_index(n, total):
    """np.pad(mode='reflect') source index for positions 0..total-1 (A4)."""
    i = np.arange(total)
    if n == 1:
        return np.zeros(total, dtype=np.int64)
    period = 2 * (n - 1)
    i = i % period
    return np.where(i < n, i, period - i)


def pad_reflect8(plane):
    """/root/reference/engines/block_processor.py:7-16 with block_size 8."""
    h, w = plane.shape
    Hp, Wp = -(-h // 8) * 8, -(-w // 8) * 8
    if (Hp, Wp) == (h, w):
        return plane
    return plane[reflect_index(h, Hp)][:, reflect_index(w, Wp)]


def _dct8(x, f):
    """A5 - one 8-point DCT-II exactly as ducc0 executes it. x: list of 8 arrays."""
    c = list(x)
    c[0] = c[0] * 2.0
    c[7] = c[7] * 2.0
    for k in (1, 3, 5):
        a, b = c[k + 1], c[k]
        c[k + 1] = a - b
        c[k] = a + b
    h = [None] * 8
    h[0] = c[0] + c[7]
    h[4] = c[0] - c[7]
    h[3] = 2.0 * c[3]
    h[7] = -2.0 * c[4]
    h[1] = c[1] + c[5]
    tr = c[1] - c[5]
    ti = c[2] + c[6]
    h[2] = c[2] - c[6]
    h[6] = (WR * ti) + (WI * tr)
    h[5] = (WR * tr) - (WI * ti)
    r = [None] * 8
    for k in (0, 1):
        b = 4 * k
        u = h[b] + h[b + 3]
        v = h[b] - h[b + 3]
        p = 2.0 * h[b + 1]
        q = 2.0 * h[b + 2]
        r[k] = u + p
        r[k + 4] = u - p
        r[k + 6] = v + q
        r[k + 2] = v - q
    if f != 1.0:
        r = [ri * f for ri in r]
    X = [None] * 8
    X[0] = r[0]
    for k, kc in ((1, 7), (2, 6), (3, 5)):
        t1 = (TW[k - 1] * r[kc]) + (TW[kc - 1] * r[k])
        t2 = (TW[k - 1] * r[k]) - (TW[kc - 1] * r[kc])
        X[k] = 0.5 * (t1 + t2)
        X[kc] = 0.5 * (t1 - t2)
    X[4] = r[4] * TW[3]
    X[0] = X[0] * HALF_S2
    return X


def _idct8(X, f):
    """A6 - one 8-point DCT-III exactly as ducc0 executes it."""
    c = list(X)
    c[0] = c[0] * S2
    for k, kc in ((1, 7), (2, 6), (3, 5)):
        t1 = c[k] + c[kc]
        t2 = c[k] - c[kc]
        c[k] = (TW[k - 1] * t2) + (TW[kc - 1] * t1)
        c[kc] = (TW[k - 1] * t1) - (TW[kc - 1] * t2)
    c[4] = c[4] * (2.0 * TW[3])
    g = [None] * 8
    for k in (0, 1):
        x0, x1, x2, x3 = c[k], c[k + 2], c[k + 4], c[k + 6]
        tr1 = x3 + x1
        g[4 * k + 2] = x3 - x1
        tr2 = x0 + x2
        g[4 * k + 1] = x0 - x2
        g[4 * k] = tr2 + tr1
        g[4 * k + 3] = tr2 - tr1
    r = [None] * 8
    r[0] = g[0] + g[4]
    r[7] = g[0] - g[4]
    r[4] = -g[7]
    r[3] = g[3]
    tr = (WR * g[5]) + (WI * g[6])
    ti = (WR * g[6]) - (WI * g[5])
    r[1] = g[1] + tr
    r[5] = g[1] - tr
    r[2] = ti + g[2]
    r[6] = ti - g[2]
    if f != 1.0:
        r = [ri * f for ri in r]
    for k in (1, 3, 5):
        a, b = r[k], r[k + 1]
        r[k] = a - b
        r[k + 1] = a + b
    return r


def dct2_blocks(blocks):
    """2-D DCT-II (ortho) of (..., 8, 8) blocks: axis -2 first, then axis -1."""
    cols = _dct8([blocks[..., i, :] for i in range(8)], 1.0 / 16.0)
    t = np.stack(cols, axis=-2)
    rows = _dct8([t[..., :, j] for j in range(8)], 1.0)
    return np.stack(rows, axis=-1)


def idct2_blocks(coefs):
    """2-D DCT-III (ortho) of (..., 8, 8) blocks: axis -2 first, then axis -1."""
    cols = _idct8([coefs[..., i, :] for i in range(8)], 1.0 / 16.0)
    t = np.stack(cols, axis=-2)
    rows = _idct8([t[..., :, j] for j in range(8)], 1.0)
    return np.stack(rows, axis=-1)


def to_blocks(padded):
    Hp, Wp = padded.shape
    return padded.reshape(Hp // 8, 8, Wp // 8, 8).transpose(0, 2, 1, 3)


def from_blocks(blocks):
    nby, nbx = blocks.shape[:2]
    return blocks.transpose(0, 2, 1, 3).reshape(nby * 8, nbx * 8)


def upsample_linear(src, H, W, ipp=True):
    """A8 - cv2.resize(f64, (W,H), INTER_LINEAR).  ``ipp=True`` is the IPP kernel
    (default in the wheel); ``ipp=False`` is OpenCV's own C++ path."""
    h, w = src.shape

    def taps(n_dst, n_src):
        # IPP computes the source coordinate with one FMA; for the exact 2x factor of even
        # sizes the product is exact and the FMA is invisible  [verified against cv2 4.13.0 /
        # IPP 2022.2 on 300 random odd/even sizes, 0 mismatches]
        x = np.arange(n_dst, dtype=np.float64)
        f = _fma(x + 0.5, np.full(n_dst, n_src / n_dst), np.full(n_dst, -0.5))
        s = np.floor(f)
        f = f - s
        s = s.astype(np.int64)
        i0 = np.clip(s, 0, n_src - 1)
        i1 = np.clip(s + 1, 0, n_src - 1)
        return i0, i1, f

    x0, x1, fx = taps(W, w)
    y0, y1, fy = taps(H, h)
    if ipp:
        if w == W:
            t = src
        else:
            t = _fma(src[:, x1] - src[:, x0], fx[None, :], src[:, x0])
        if h == H:
            return np.array(t, copy=True)
        return _fma(t[y1] - t[y0], fy[:, None], t[y0])
    raise NotImplementedError("non-IPP path is documented in SURVEY A8, not restated")


def ycbcr_to_rgb(Y, Cb, Cr):
    """A9 - /root/reference/engines/color_space.py:17-24 (clip inside)."""
    R = Y + 1.402 * (Cr - 128.0)
    G = Y - 0.344136 * (Cb - 128.0) - 0.714136 * (Cr - 128.0)
    B = Y + 1.772 * (Cb - 128.0)
    return np.clip(np.stack([R, G, B], axis=-1), 0, 255)


def bit_length_sum(coeffs):
    """Exact integer version of /root/reference/utils/metrics.py:75-79."""
    nz = coeffs[coeffs != 0]
    mag = np.abs(nz.astype(np.int64))
    bl = np.floor(np.log2(mag)).astype(np.int64) + 1      # == bit_length for 1..2047
    return int(nz.size), int(6 * nz.size + np.sum(bl + 1))


def bitrate_reference_arithmetic(coeffs_i16, shape):
    """/root/reference/utils/metrics.py:51-92 verbatim arithmetic incl. the float32
    accumulation under NumPy 2 (SURVEY §8a 'Quirk')."""
    h, w = shape
    num_pixels = h * w
    original_bits = num_pixels * 3 * 8
    num_blocks = (-(-h // 8)) * (-(-w // 8))
    block_overhead_bits = num_blocks * 2
    mask = coeffs_i16 != 0
    nz = coeffs_i16[mask]
    if len(nz) > 0:
        position_bits = 6 * len(nz)
        magnitudes = np.abs(nz)
        magnitude_bits = np.sum(np.ceil(np.log2(magnitudes + 1)) + 1)
        coeff_bits = position_bits + magnitude_bits
    else:
        coeff_bits = 0
    estimated_bits = block_overhead_bits + coeff_bits
    return {
        "estimated_bits": int(estimated_bits),
        "bpp": float(estimated_bits / num_pixels),
        "compression_ratio": float(original_bits / max(estimated_bits, 1)),
        "nonzero_count": int(np.sum(mask)),
        "total_coeffs": int(coeffs_i16.size),
    }


def psnr_ssim(orig_u8, recon_u8):
    """/root/reference/utils/metrics.py:9-28."""
    psnr_rgb = _sk.peak_signal_noise_ratio(orig_u8, recon_u8, data_range=255)
    ssim_rgb = _sk.structural_similarity(orig_u8, recon_u8, channel_axis=2, data_range=255)
    oy = 0.299 * orig_u8[:, :, 0] + 0.587 * orig_u8[:, :, 1] + 0.114 * orig_u8[:, :, 2]
    ry = 0.299 * recon_u8[:, :, 0] + 0.587 * recon_u8[:, :, 1] + 0.114 * recon_u8[:, :, 2]
    psnr_y = _sk.peak_signal_noise_ratio(oy, ry, data_range=255)
    ssim_y = _sk.structural_similarity(oy, ry, data_range=255)
    return {"psnr_rgb": float(psnr_rgb), "ssim_rgb": float(ssim_rgb),
            "psnr_y": float(psnr_y), "ssim_y": float(ssim_y)}


# ---------------------------------------------------------------------------
# the round trip
# ---------------------------------------------------------------------------

def encode_plane(plane, Q):
    """pad -> blocks -> (-128) -> DCT -> quantise (A4, A5, A7)."""
    padded = pad_reflect8(plane)
    blocks = to_blocks(padded)
    dct = dct2_blocks(blocks - 128.0)
    q = np.round(dct / Q).astype(np.int16)
    return padded, blocks, dct, q


def decode_plane(q, Q, shape):
    """dequantise -> IDCT -> +128 -> clip -> merge -> crop (A6, A7)."""
    deq = q.astype(np.float64) * Q
    rec = np.clip(idct2_blocks(deq) + 128.0, 0, 255)
    h, w = shape
    return from_blocks(rec)[:h, :w], deq, rec


def compress_reconstruct(image_rgb, quality=50, mode="4:2:0", prefilter=False,
                         selected_block_idx=(0, 0), want_metrics=True,
                         want_maps=True):
    """The whole path; returns a plain dict (no dataclasses: this is the checker)."""
    if mode not in MODES:
        raise ValueError(f"Unknown subsampling mode: {mode}")
    image_rgb = np.asarray(image_rgb)
    H, W = image_rgb.shape[:2]
    img_f = image_rgb.astype(np.float64)
    Y, Cb, Cr = rgb_to_ycbcr(img_f)
    if mode == "4:4:4":
        Cb_s, Cr_s = Cb, Cr
    else:
        if prefilter:
            Cb, Cr = gaussian_blur_3x3(Cb), gaussian_blur_3x3(Cr)
        Cb_s, Cr_s = decimate_area(Cb, mode), decimate_area(Cr, mode)
    Q = scale_quant_matrix(quality)

    enc = [encode_plane(p, Q) for p in (Y, Cb_s, Cr_s)]
    coeffs = np.concatenate([e[3].reshape(-1) for e in enc])
    dec = [decode_plane(e[3], Q, p.shape) for e, p in zip(enc, (Y, Cb_s, Cr_s))]
    Y_r, Cb_r, Cr_r = (d[0] for d in dec)
    if mode != "4:4:4":
        Cb_r = upsample_linear(Cb_r, H, W)
        Cr_r = upsample_linear(Cr_r, H, W)
    rgb_f = ycbcr_to_rgb(Y_r, Cb_r, Cr_r)
    recon = np.clip(rgb_f, 0, 255).astype(np.uint8)

    out = {"reconstructed_image": recon, "all_quantized_coeffs": coeffs,
           "Q_matrix": Q, "shape": (H, W)}
    nnz, bits_coef = bit_length_sum(coeffs)
    nblk = (-(-H // 8)) * (-(-W // 8))
    out["exact_bits"] = 2 * nblk + bits_coef
    out["nonzero_coeffs"] = nnz
    out["total_coeffs"] = int(coeffs.size)
    out.update({k: v for k, v in bitrate_reference_arithmetic(coeffs, (H, W)).items()
                if k in ("estimated_bits", "bpp", "compression_ratio")})
    if want_metrics:
        out.update(psnr_ssim(image_rgb, recon))
    if want_maps:
        out["error_map_y"] = np.abs(Y - Y_r)
        out["error_map_rgb"] = np.mean(np.abs(img_f - rgb_f), axis=2)
        out["quantized_histogram"] = np.histogram(coeffs, bins=50, range=(-100, 100))[0]
    # selected block (pipeline.py:126-151) - Y channel only
    padded_y, blocks_y, dct_y, q_y = enc[0]
    nby, nbx = blocks_y.shape[:2]
    br, bc = selected_block_idx
    t = br * nbx + bc
    sel = None
    if 0 <= t < nby * nbx:
        i, j = divmod(t, nbx)
        sel = {
            "original": blocks_y[i, j].copy(),
            "shifted": blocks_y[i, j] - 128.0,
            "dct": dct_y[i, j].copy(),
            "quantized": q_y[i, j].copy(),
            "dequantized": dec[0][1][i, j].copy(),
            "reconstructed": dec[0][2][i, j].copy(),
        }
    out["selected_block"] = sel
    out["selected_block_idx"] = selected_block_idx
    return out


# ------------------------------------------------------------------------------------------
# Preview downscale in front of the round trip
# ------------------------------------------------------------------------------------------
def preview_size(h, w, target_w, target_h):
    """/root/reference/gui/compression_tab.py:538-547 (_update_preview_image)."""
    if w <= target_w and h <= target_h:
        return h, w
    scale = min(target_w / w, target_h / h)
    return int(h * scale), int(w * scale)


def resize_area_u8(src, dh, dw):
    """cv2.resize(uint8 H x W x C, (dw, dh), interpolation=cv2.INTER_AREA) for shrinking
    (/root/reference/gui/compression_tab.py:549-552).  OpenCV 4.13 imgproc/resize.cpp:
    integer factors on both axes -> resizeAreaFast_ (int sum; 2x2 -> (sum+2)>>2, else
    saturate_cast<uchar>(sum * (1.f/area))); otherwise ResizeArea_Invoker<uchar, float>
    with the taps of area_tab, fp32 accumulation without FMA.
    [verified against cv2 4.13.0: tests/test_preview_cpu.py]"""
    sh, sw, cn = src.shape
    sx, sy = sw / dw, sh / dh
    isx, isy = int(np.rint(sx)), int(np.rint(sy))
    eps = np.finfo(np.float64).eps
    if abs(sx - isx) < eps and abs(sy - isy) < eps:
        a = src[:dh * isy, :dw * isx].reshape(dh, isy, dw, isx, cn).astype(np.int64).sum(axis=(1, 3))
        if isx == 2 and isy == 2:
            return ((a + 2) >> 2).astype(np.uint8)
        v = a.astype(np.float32) * (np.float32(1.0) / np.float32(isx * isy))
        return np.clip(np.rint(v), 0, 255).astype(np.uint8)
    f32 = np.float32
    S = src.astype(f32)
    buf = np.zeros((sh, dw, cn), dtype=f32)
    for dx, s_, a in area_tab(sw, dw):
        buf[:, dx, :] = buf[:, dx, :] + S[:, s_, :] * f32(a)
    out = np.zeros((dh, dw, cn), dtype=f32)
    first = np.ones(dh, dtype=bool)
    for dy, s_, b in area_tab(sh, dh):
        if first[dy]:
            out[dy] = f32(b) * buf[s_]
            first[dy] = False
        else:
            out[dy] = out[dy] + f32(b) * buf[s_]
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def make_preview(image, target_w, target_h):
    """/root/reference/gui/compression_tab.py:532-552."""
    h, w = image.shape[:2]
    nh, nw = preview_size(h, w, target_w, target_h)
    if (nh, nw) == (h, w):
        return image.copy()
    return resize_area_u8(image, nh, nw)
