"""Synthetic test-image generators (host side, NumPy).

Same names, arguments and pixel values as the reference's ``utils/test_images.py``
(checkerboard :6-19, stripes :22-32, gradient :35-48, text edges :51-77, chroma
stripes :80-102, photo :105-162, ``generate_demo_image`` :165-178); written as array
expressions instead of per-pixel Python loops.  ``tests/test_inputs.py`` checks them
byte-for-byte against the reference's generators where the reference is present.
"""

import numpy as np


def generate_colored_checkerboard(size: int = 512) -> np.ndarray:
    """High-contrast 32-px checkerboard of (30,30,30) / (220,220,220)."""
    idx = np.arange(size) // 32
    odd = ((idx[:, None] + idx[None, :]) % 2).astype(bool)
    img = np.where(odd[..., None], np.uint8(220), np.uint8(30))
    return np.ascontiguousarray(np.broadcast_to(img, (size, size, 3)), dtype=np.uint8)


def generate_thin_stripes(size: int = 512, stripe_width: int = 4) -> np.ndarray:
    """Fine vertical stripes alternating (200,60,60) and (60,180,200)."""
    even = ((np.arange(size) // stripe_width) % 2 == 0)
    row = np.where(even[:, None], np.array([200, 60, 60], np.uint8),
                   np.array([60, 180, 200], np.uint8))
    return np.ascontiguousarray(np.broadcast_to(row[None], (size, size, 3)), dtype=np.uint8)


def generate_gradient(size: int = 512) -> np.ndarray:
    """Smooth diagonal gradient (computed in fp64, stored through float32 like the
    reference, then truncated to uint8)."""
    ij = np.arange(size)
    t = (ij[:, None] + ij[None, :]) / (2 * size - 2)
    img = np.stack([40 + t * 180, 60 + t * 140, 120 + t * 100], axis=-1).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def _bar_runs(first: int, unit: int, gap: int):
    """Start offsets and thicknesses of the four bars (unit, unit/2, unit/4, 2 px)."""
    runs, pos = [], first
    for t in (unit, unit // 2, unit // 4, 2):
        runs.append((pos, t))
        pos += t + gap
    return runs


def generate_text_edges(size: int = 512) -> np.ndarray:
    """Dark (25) bars of shrinking thickness - four horizontal in the upper half, four
    vertical in the lower half - and a 3-px diagonal, on a light (245) background."""
    m = size // 10                      # margin
    dark = np.zeros((size, size), dtype=bool)
    for y, t in _bar_runs(m, size // 16, m // 2):
        dark[y:y + t, m:size - m] = True
    for x, t in _bar_runs(m, size // 16, m // 2):
        dark[size // 2 + m:size - m, x:x + t] = True
    i = np.arange(size // 4)
    yy, xx = size // 2 + m + i, size // 2 + i
    ok = (yy < size - m) & (xx < size - m)
    for dy in range(3):
        for dx in range(3):
            r, c = yy[ok] + dy, xx[ok] + dx
            keep = (r < size) & (c < size)
            dark[r[keep], c[keep]] = True
    img = np.where(dark[..., None], np.uint8(25), np.uint8(245))
    return np.ascontiguousarray(np.broadcast_to(img, (size, size, 3)), dtype=np.uint8)


_CHROMA_BARS = np.array([
    [180, 40, 40], [40, 160, 40], [40, 80, 180], [180, 180, 40],
    [180, 40, 180], [40, 180, 180], [200, 120, 40], [120, 40, 180]], dtype=np.uint8)


def generate_chroma_stripes(size: int = 512) -> np.ndarray:
    """Eight saturated vertical colour bars."""
    stripe_width = size // len(_CHROMA_BARS)
    if stripe_width == 0:
        return np.zeros((size, size, 3), dtype=np.uint8)
    bar = np.minimum(np.arange(size) // stripe_width, len(_CHROMA_BARS) - 1)
    return np.ascontiguousarray(np.broadcast_to(_CHROMA_BARS[bar][None], (size, size, 3)))


def generate_photo(size: int = 512) -> np.ndarray:
    """Synthetic landscape: sky gradient, mountains, noisy ground, sun glow.

    Consumes the legacy global NumPy RNG exactly like the reference
    (``np.random.seed(123)``, 4 phases, then one draw per ground pixel in raster
    order) so the bytes are identical.
    """
    img = np.zeros((size, size, 3), dtype=np.float32)
    horizon = int(size * 0.45)
    mountain_base = int(size * 0.55)

    t = np.arange(horizon) / horizon
    img[:horizon] = np.stack([180 - t * 60, 210 - t * 80, 240 - t * 40], axis=-1)[:, None, :]

    np.random.seed(123)
    heights = np.zeros(size)
    for freq in [8, 16, 32, 64]:
        phase = np.random.rand() * 2 * np.pi
        amplitude = (size * 0.15) / (freq / 8)
        heights += amplitude * np.sin(np.linspace(0, freq * np.pi, size) + phase)
    heights = heights - heights.min()
    heights = heights / heights.max() * (mountain_base - horizon - 20)

    peak = (horizon + 20 + heights).astype(np.int64)            # int() truncation, values > 0
    rows = np.arange(horizon, mountain_base)
    if rows.size:
        below = rows[:, None] < peak[None, :]
        depth = (rows[:, None] - horizon) / (peak[None, :] - horizon)
        rock = np.stack([70 + depth * 30, 80 + depth * 20, 100 + depth * 10], axis=-1)
        flat = np.broadcast_to(np.array([90.0, 95.0, 85.0]), rock.shape)
        img[horizon:mountain_base] = np.where(below[..., None], rock, flat)

    n_ground = size - mountain_base
    if n_ground > 0:
        tg = (np.arange(mountain_base, size) - mountain_base) / (size - mountain_base)
        noise = np.random.rand(n_ground * size).reshape(n_ground, size) * 15 - 7.5
        tg = tg[:, None]
        img[mountain_base:] = np.stack([60 + tg * 40 + noise, 100 + tg * 30 + noise,
                                        50 + tg * 20 + noise], axis=-1)

    sun_x, sun_y = size // 4, size // 6
    sun_radius = size // 10
    i0, i1 = max(0, sun_y - sun_radius * 2), min(horizon, sun_y + sun_radius * 2)
    j0, j1 = max(0, sun_x - sun_radius * 2), min(size, sun_x + sun_radius * 2)
    if i1 > i0 and j1 > j0:
        ii = np.arange(i0, i1)[:, None]
        jj = np.arange(j0, j1)[None, :]
        dist = np.sqrt((ii - sun_y) ** 2 + (jj - sun_x) ** 2)
        lim = sun_radius * 1.5
        inside = dist < lim
        with np.errstate(divide="ignore", invalid="ignore"):
            glow = np.maximum(0, 1 - (dist / lim) ** 2)
        region = img[i0:i1, j0:j1]
        # reference: float32 pixel * np.float64 scalar -> fp64 (NEP 50), + int64 array * glow * 0.7
        blended = region.astype(np.float64) * (1 - glow * 0.7)[..., None] + \
            (np.array([255, 240, 200]) * glow[..., None]) * 0.7
        img[i0:i1, j0:j1] = np.where(inside[..., None], blended.astype(np.float32), region)

    return np.clip(img, 0, 255).astype(np.uint8)


_DEMO = {"photo": generate_photo, "text_edges": generate_text_edges, "gradient": generate_gradient,
         "checkerboard": generate_colored_checkerboard, "chroma_stripes": generate_chroma_stripes}


def generate_demo_image(key: str):
    """512-px demo image by key; ``None`` for a key the GUI does not know."""
    make = _DEMO.get(key)
    return make(512) if make else None
