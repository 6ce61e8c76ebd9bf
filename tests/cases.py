"""Named parity cases shared by the golden generator, the CPU tests and the GPU tests.

Inputs are produced by this repo's own generators / seeded NumPy RNGs so that they
exist on the GPU box (where /root/reference does not).  The golden file stores the
SHA-256 of every input, so a drift in a generator is caught before any parity claim.

Sources: SURVEY.md §8c (fixture list), §8d (concrete inputs of BASELINE.json's
configs cfg1..cfg5) and the reference's tests (tests/test_pipeline.py,
tests/test_subsampling.py) for the 64x64 / 256x256 shapes.
"""

import hashlib
from dataclasses import dataclass
from typing import Callable, Tuple

import numpy as np

from jpeg_dsp_studio_b200.utils import test_images as TI


def rand_rgb(seed: int, h: int, w: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def photo_tiled(h: int, w: int) -> np.ndarray:
    """generate_photo(512) tiled to (h, w): natural coefficient statistics at size."""
    p = TI.generate_photo(512)
    reps = (-(-h // 512), -(-w // 512), 1)
    return np.ascontiguousarray(np.tile(p, reps)[:h, :w])


@dataclass(frozen=True)
class Case:
    name: str
    make: Callable[[], np.ndarray]
    quality: int
    mode: str
    prefilter: bool
    sel: Tuple[int, int] = (0, 0)
    big: bool = False           # skipped by the quick CPU suite; run on the GPU box

    def image(self) -> np.ndarray:
        return self.make()


CASES = [
    # --- BASELINE.json configs -------------------------------------------------
    Case("cfg1_checker512_q10_420", lambda: TI.generate_colored_checkerboard(512), 10, "4:2:0", False),
    Case("cfg2_rand1080p_q50_444", lambda: rand_rgb(2, 1080, 1920), 50, "4:4:4", False, big=True),
    Case("cfg3_rand4k_q75_420_pf", lambda: rand_rgb(3, 2160, 3840), 75, "4:2:0", True, big=True),
    Case("cfg4_rand4k_q50_420", lambda: rand_rgb(4, 2160, 3840), 50, "4:2:0", False, big=True),
    Case("cfg5_frame0_1080p_q30_422", lambda: rand_rgb(5000, 1080, 1920), 30, "4:2:2", False, big=True),
    # --- SURVEY §8c fixtures ---------------------------------------------------
    Case("checker256_q50_420_pf", lambda: TI.generate_colored_checkerboard(256), 50, "4:2:0", True, (3, 5)),
    Case("checker256_q50_420", lambda: TI.generate_colored_checkerboard(256), 50, "4:2:0", False),
    Case("stripes256w2_q50_422", lambda: TI.generate_thin_stripes(256, 2), 50, "4:2:2", False),
    Case("stripes256w2_q50_422_pf", lambda: TI.generate_thin_stripes(256, 2), 50, "4:2:2", True),
    Case("stripes512w4_q35_420", lambda: TI.generate_thin_stripes(512, 4), 35, "4:2:0", False, (10, 63)),
    Case("chroma512_q10_420", lambda: TI.generate_chroma_stripes(512), 10, "4:2:0", False),
    Case("gradient512_q90_444", lambda: TI.generate_gradient(512), 90, "4:4:4", False, (63, 63)),
    Case("text512_q60_422_pf", lambda: TI.generate_text_edges(512), 60, "4:2:2", True, (7, 7)),
    Case("photo512_q75_420_pf", lambda: TI.generate_photo(512), 75, "4:2:0", True, (40, 1)),
    Case("photo512_q20_444", lambda: TI.generate_photo(512), 20, "4:4:4", False),
    # --- padding (reflect) and ragged shapes -------------------------------------
    Case("rand250x334_q50_420_pf", lambda: rand_rgb(0, 250, 334), 50, "4:2:0", True, (2, 3)),
    Case("rand250x334_q50_422", lambda: rand_rgb(0, 250, 334), 50, "4:2:2", False, (31, 41)),
    Case("rand250x334_q90_444", lambda: rand_rgb(0, 250, 334), 90, "4:4:4", False, (-1, 3)),
    Case("rand18x22_q50_420", lambda: rand_rgb(7, 18, 22), 50, "4:2:0", False, (0, 9)),
    Case("rand8x8_q50_444", lambda: rand_rgb(8, 8, 8), 50, "4:4:4", False),
    Case("rand64x64_q100_444", lambda: rand_rgb(9, 64, 64), 100, "4:4:4", False),
    Case("rand64x64_q1_420", lambda: rand_rgb(9, 64, 64), 1, "4:2:0", False),
    Case("rand72x40_q99_422_pf", lambda: rand_rgb(10, 72, 40), 99, "4:2:2", True, (8, 4)),
    Case("rand30x46_q25_420_pf", lambda: rand_rgb(12, 30, 46), 25, "4:2:0", True),
    # --- odd sizes with subsampling: cv2's fractional INTER_AREA taps and IPP's non-2x
    #     bilinear factors (engines/color_space.py:44-49, 64-65) ----------------------------
    Case("rand33x47_q50_420", lambda: rand_rgb(20, 33, 47), 50, "4:2:0", False, (1, 2)),
    Case("rand33x47_q50_420_pf", lambda: rand_rgb(20, 33, 47), 50, "4:2:0", True),
    Case("rand47x33_q30_422_pf", lambda: rand_rgb(21, 47, 33), 30, "4:2:2", True),
    Case("rand48x33_q80_420", lambda: rand_rgb(22, 48, 33), 80, "4:2:0", False),
    Case("rand33x48_q20_420_pf", lambda: rand_rgb(23, 33, 48), 20, "4:2:0", True),
    Case("rand99x101_q10_422", lambda: rand_rgb(24, 99, 101), 10, "4:2:2", False),
    Case("rand9x7_q50_420_pf", lambda: rand_rgb(25, 9, 7), 50, "4:2:0", True),
    Case("rand251x335_q75_420_pf", lambda: rand_rgb(26, 251, 335), 75, "4:2:0", True, (30, 41)),
    Case("photo301x401_q60_420", lambda: np.ascontiguousarray(TI.generate_photo(512)[:301, :401]), 60, "4:2:0", False),
    Case("preview853x1280_q50_420", lambda: photo_tiled(853, 1280), 50, "4:2:0", False, big=True),
    # flat image: identical round trip -> PSNR inf, SSIM 1 (SURVEY §8a error list)
    Case("flat64_q50_444", lambda: np.full((64, 64, 3), 128, np.uint8), 50, "4:4:4", False),
    Case("phototile1080p_q50_420", lambda: photo_tiled(1080, 1920), 50, "4:2:0", False, big=True),
]

BY_NAME = {c.name: c for c in CASES}

#: quality sweep fixtures (cfg4 at test size): every Q in 1..100, four mode/prefilter
#: combinations, on one small random frame
SWEEP_IMAGE = ("rand160x224", lambda: rand_rgb(11, 160, 224))
SWEEP_COMBOS = [("4:2:0", False), ("4:2:0", True), ("4:2:2", False), ("4:4:4", False)]
SWEEP_QUALITIES = list(range(1, 101))


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# --- chroma-aliasing demo (gui/dialogs/aliasing_demo_dialog.py) -----------
@dataclass(frozen=True)
class AliasingCase:
    name: str
    make: Callable[[], np.ndarray]
    quality: int

    def image(self) -> np.ndarray:
        return self.make()


def _alias_patterns():
    from jpeg_dsp_studio_b200.engines import aliasing_demo as AD
    return AD


ALIASING_CASES = [
    # the dialog's three synthetic patterns at its default size and quality (:20-66, Q=50)
    AliasingCase("stripes256_q50", lambda: _alias_patterns().generate_equiluminance_stripes(256), 50),
    AliasingCase("chroma_checker256_q50", lambda: _alias_patterns().generate_chroma_checkerboard(256), 50),
    AliasingCase("checker1px256_q50", lambda: _alias_patterns().generate_1px_checkerboard(256), 50),
    AliasingCase("checker1px64_q10", lambda: _alias_patterns().generate_1px_checkerboard(64), 10),
    # loaded images / ROI crops of any size (:398-434): block-aligned, ragged and odd
    AliasingCase("rand96x128_q50", lambda: rand_rgb(31, 96, 128), 50),
    AliasingCase("rand57x75_q35", lambda: rand_rgb(32, 57, 75), 35),
    AliasingCase("rand33x70_q90", lambda: rand_rgb(33, 33, 70), 90),
    AliasingCase("rand40x9_q50", lambda: rand_rgb(34, 40, 9), 50),
    AliasingCase("photo200x264_q60", lambda: np.ascontiguousarray(TI.generate_photo(512)[100:300, 40:304]), 60),
    AliasingCase("photo512_q25", lambda: TI.generate_photo(512), 25),
]
