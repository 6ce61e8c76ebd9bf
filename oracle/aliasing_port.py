"""NumPy restatement of the reference's chroma-aliasing demo (TEST INFRASTRUCTURE).

Follows ``/root/reference/gui/dialogs/aliasing_demo_dialog.py``:
  * pattern generators                                   :20-66
  * ``compute_metrics``                                  :69-83
  * ``AliasingDemoWorker._process_with_explicit_subsample``  :125-160
  * ``AliasingDemoWorker._compute_difference``           :162-166
  * ``AliasingDemoWorker.run``                           :98-123

The demo is a *different arithmetic* from the hot path:
OpenCV float32 kernels - ``cvtColor(RGB2YCrCb)`` on a float image, ``GaussianBlur(5x5, 0.8)``,
``[::2, ::2]`` decimation, ``resize(INTER_LINEAR)``, ``cvtColor(YCrCb2RGB)`` - followed by the
hot path at 4:4:4.  The operation order of those kernels (OpenCV 4.13.0, AVX2/AVX-512 build)
was probed in the build container and is pinned by tests/test_oracle_vs_reference.py against
live OpenCV and against the reference's own worker:

  F1  RGB2YCrCb, float32, per image row:  columns below 8*floor(W/8) (vector body)
          Y = fma(R, .299f, fma(G, .587f, B*.114f))
      the remaining columns (scalar tail)
          Y = fma(B, .114f, fma(R, .299f, G*.587f))
      everywhere  Cr = fma(R - Y, .713f, .5f),  Cb = fma(B - Y, .564f, .5f)
      (delta is 0.5 for float images although the data is 0..255 - the reference's quirk).
  F2  GaussianBlur(5x5, sigma .8), float32, BORDER_REFLECT_101, kernel = float32 of
      cv2.getGaussianKernel(5, .8): row pass  fma(m2+p2, k0, fma(c, k2, (m1+p1)*k1)),
      except the last column of an odd width, which is the un-fused
      (c*k2 + (m1+p1)*k1) + (m2+p2)*k0;  column pass
      fma(m2+p2, k0, fma(m1+p1, k1, c*k2)) for columns below 8*floor(W/8), the same un-fused
      form beyond.
  F3  resize(INTER_LINEAR), float32: source coordinate in fp64 with one FMA
      (fma(x+.5, n_src/n_dst, -.5), as A8'), weight cast to float32, then
      fma(b-a, w, a) horizontally and the same vertically (float32 FMAs).
  F4  YCrCb2RGB, float32, all columns:  cr = Cr-.5f, cb = Cb-.5f,
          R = fma(cr, 1.403f, Y),  G = fma(cr, -.714f, fma(cb, -.344f, Y)),  B = fma(cb, 1.773f, Y).
  F5  RGB2YCrCb on uint8 (compute_metrics): Y = (4899 R + 9617 G + 1868 B + 8192) >> 14
      (checked on all 2^24 colours).

Parity status: PINNED (same evidence as numpy_port; SSIM to skimage's published algorithm).
"""

import numpy as np

from . import numpy_port as P
from . import skimage_standin as _sk

f32 = np.float32

# cv2.getGaussianKernel(5, 0.8, CV_32F)
K0 = f32(float.fromhex("0x1.674b98p-6"))
K1 = f32(float.fromhex("0x1.d3fe2ep-3"))
K2 = f32(float.fromhex("0x1.ff1860p-2"))


def fma32(a, b, c):
    """Correctly rounded float32 fused multiply-add: the product of two float32 is exact
    in fp64; the sum is rounded to ODD in fp64 (TwoSum error term as the sticky bit), so the
    final rounding to float32 is the single rounding of the exact value."""
    a = np.asarray(a, f32).astype(np.float64)
    b = np.asarray(b, f32).astype(np.float64)
    c = np.asarray(c, f32).astype(np.float64)
    a, b, c = np.broadcast_arrays(a, b, c)
    p = a * b
    s = p + c
    bb = s - p
    t = (p - (s - bb)) + (c - bb)
    need = (t != 0) & ((s.view(np.int64) & 1) == 0)
    if np.any(need):
        s = s.copy()
        s[need] = np.nextafter(s[need], np.where(t[need] > 0, np.inf, -np.inf))
    return s.astype(f32)


def _mul(a, b):
    return (np.asarray(a, f32) * np.asarray(b, f32)).astype(f32)


def _add(a, b):
    return (np.asarray(a, f32) + np.asarray(b, f32)).astype(f32)


def _sub(a, b):
    return (np.asarray(a, f32) - np.asarray(b, f32)).astype(f32)


# ---------------------------------------------------------------------------
# pattern generators (aliasing_demo_dialog.py:20-66), vectorised
# ---------------------------------------------------------------------------
def generate_equiluminance_stripes(size=256):
    img = np.empty((size, size, 3), dtype=np.uint8)
    img[:, 0::2] = (220, 40, 60)
    img[:, 1::2] = (30, 220, 210)
    return img


def generate_chroma_checkerboard(size=256):
    i, j = np.indices((size, size))
    m = ((i // 2) + (j // 2)) % 2 == 0
    return np.where(m[..., None], np.array([230, 50, 60], np.uint8), np.array([50, 220, 220], np.uint8))


def generate_1px_checkerboard(size=256):
    i, j = np.indices((size, size))
    m = (i + j) % 2 == 0
    return np.where(m[..., None], np.array([240, 40, 50], np.uint8), np.array([40, 240, 230], np.uint8))


# ---------------------------------------------------------------------------
# float32 OpenCV kernels
# ---------------------------------------------------------------------------
def rgb_to_ycrcb_f32(img_f32):
    """F1.  Returns Y, Cr, Cb (float32 planes)."""
    R, G, B = (np.ascontiguousarray(img_f32[..., k], dtype=f32) for k in range(3))
    W = R.shape[1]
    W8 = 8 * (W // 8)
    c0, c1, c2 = f32(0.299), f32(0.587), f32(0.114)
    Y = np.empty_like(R)
    Y[:, :W8] = fma32(R[:, :W8], c0, fma32(G[:, :W8], c1, _mul(B[:, :W8], c2)))
    if W8 < W:
        Y[:, W8:] = fma32(B[:, W8:], c2, fma32(R[:, W8:], c0, _mul(G[:, W8:], c1)))
    Cr = fma32(_sub(R, Y), f32(0.713), f32(0.5))
    Cb = fma32(_sub(B, Y), f32(0.564), f32(0.5))
    return Y, Cr, Cb


def _reflect101(i, n):
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def gaussian_blur_5x5_f32(x):
    """F2 - cv2.GaussianBlur(x, (5, 5), 0.8) on a float32 plane (both sides >= 3)."""
    x = np.asarray(x, f32)
    H, W = x.shape
    W8 = 8 * (W // 8)
    cols = np.arange(W)
    m2, m1, c, p1, p2 = (x[:, _reflect101(cols + d, W)] for d in (-2, -1, 0, 1, 2))
    s2, s1 = _add(m2, p2), _add(m1, p1)
    r = fma32(s2, K0, fma32(c, K2, _mul(s1, K1)))
    if W % 2 == 1:
        j = W - 1
        r[:, j] = _add(_add(_mul(c[:, j], K2), _mul(s1[:, j], K1)), _mul(s2[:, j], K0))
    rows = np.arange(H)
    m2, m1, c, p1, p2 = (r[_reflect101(rows + d, H), :] for d in (-2, -1, 0, 1, 2))
    s2, s1 = _add(m2, p2), _add(m1, p1)
    out = fma32(s2, K0, fma32(s1, K1, _mul(c, K2)))
    if W8 < W:
        out[:, W8:] = _add(_add(_mul(c[:, W8:], K2), _mul(s1[:, W8:], K1)), _mul(s2[:, W8:], K0))
    return out


def resize_linear_f32(src, H, W):
    """F3 - cv2.resize(src, (W, H), interpolation=cv2.INTER_LINEAR), float32, enlarging."""
    src = np.asarray(src, f32)
    h, w = src.shape

    def taps(n_dst, n_src):
        x = np.arange(n_dst, dtype=np.float64)
        f = P._fma(x + 0.5, np.full(n_dst, n_src / n_dst), np.full(n_dst, -0.5))
        s = np.floor(f)
        f = (f - s).astype(f32)
        s = s.astype(np.int64)
        return np.clip(s, 0, n_src - 1), np.clip(s + 1, 0, n_src - 1), f

    x0, x1, fx = taps(W, w)
    y0, y1, fy = taps(H, h)
    t = fma32(_sub(src[:, x1], src[:, x0]), fx[None, :], src[:, x0])
    return fma32(_sub(t[y1], t[y0]), fy[:, None], t[y0])


def ycrcb_to_rgb_f32(Y, Cr, Cb):
    """F4.  Returns H x W x 3 float32 (unclipped)."""
    cr, cb = _sub(Cr, f32(0.5)), _sub(Cb, f32(0.5))
    R = fma32(cr, f32(1.403), Y)
    G = fma32(cr, f32(-0.714), fma32(cb, f32(-0.344), Y))
    B = fma32(cb, f32(1.773), Y)
    return np.stack([R, G, B], axis=-1)


def luma_u8(img_u8):
    """F5 - Y channel of cv2.cvtColor(uint8 RGB, COLOR_RGB2YCrCb)."""
    a = img_u8.astype(np.int64)
    return ((a[..., 0] * 4899 + a[..., 1] * 9617 + a[..., 2] * 1868 + (1 << 13)) >> 14).astype(np.uint8)


# ---------------------------------------------------------------------------
# the demo
# ---------------------------------------------------------------------------
def explicit_subsample_rgb(image_u8, prefilter):
    """aliasing_demo_dialog.py:127-150 - the uint8 RGB frame handed to the hot path."""
    img = image_u8.astype(f32)
    Y, Cr, Cb = rgb_to_ycrcb_f32(img)
    h, w = Y.shape
    if prefilter:
        Cb_s = gaussian_blur_5x5_f32(Cb)[::2, ::2]
        Cr_s = gaussian_blur_5x5_f32(Cr)[::2, ::2]
    else:
        Cb_s, Cr_s = Cb[::2, ::2], Cr[::2, ::2]
    Cb_up = resize_linear_f32(Cb_s, h, w)
    Cr_up = resize_linear_f32(Cr_s, h, w)
    rgb = ycrcb_to_rgb_f32(Y, Cr_up, Cb_up)
    return np.clip(rgb, 0, 255).astype(np.uint8)


def process_with_explicit_subsample(image_u8, quality, prefilter):
    """aliasing_demo_dialog.py:125-160."""
    rgb = explicit_subsample_rgb(image_u8, prefilter)
    return P.compress_reconstruct(rgb, quality, "4:4:4", False, want_maps=False)["reconstructed_image"]


def compute_metrics(original, reconstructed):
    """aliasing_demo_dialog.py:69-83."""
    oy, ry = luma_u8(original), luma_u8(reconstructed)
    return {
        "psnr_y": _sk.peak_signal_noise_ratio(oy, ry, data_range=255),
        "ssim_y": _sk.structural_similarity(oy, ry, data_range=255),
        "psnr_rgb": _sk.peak_signal_noise_ratio(original, reconstructed, data_range=255),
        "ssim_rgb": _sk.structural_similarity(original, reconstructed, data_range=255, channel_axis=2),
    }


def compute_difference(original, reconstructed):
    """aliasing_demo_dialog.py:162-166."""
    diff = np.abs(original.astype(f32) - reconstructed.astype(f32))
    return np.clip(diff * 10, 0, 255).astype(np.uint8)


def run_demo(image_u8, quality=50):
    """aliasing_demo_dialog.py:98-123 - the dict AliasingDemoWorker emits."""
    a = process_with_explicit_subsample(image_u8, quality, False)
    b = process_with_explicit_subsample(image_u8, quality, True)
    return {
        "original": image_u8, "recon_no_pf": a, "recon_pf": b,
        "diff_no_pf": compute_difference(image_u8, a), "diff_pf": compute_difference(image_u8, b),
        "metrics_no_pf": compute_metrics(image_u8, a), "metrics_pf": compute_metrics(image_u8, b),
    }
