"""Accuracy of the fp32 strip SSIM kernel on content far from mid-grey (worst case for the centred
fp32 window sums): GPU compute_psnr_ssim vs the fp64 oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jpeg_dsp_studio_b200.utils.metrics import compute_psnr_ssim
from oracle import numpy_port as P

rng = np.random.default_rng(0)
for name, lo, hi in (("dark 0..5", 0, 6), ("bright 250..255", 250, 256), ("dark 0..1", 0, 2), ("mid 126..130", 126, 131),
                     ("full range", 0, 256)):
    a = rng.integers(lo, hi, (544, 960, 3), dtype=np.uint8)
    b = np.clip(a.astype(np.int16) + rng.integers(-1, 2, a.shape), 0, 255).astype(np.uint8)
    g = compute_psnr_ssim(a, b)
    o = P.psnr_ssim(a, b)
    print(f"{name:18s} ssim_y gpu {g['ssim_y']:.9f} oracle {o['ssim_y']:.9f} diff {g['ssim_y'] - o['ssim_y']:+.2e}   "
          f"ssim_rgb diff {g['ssim_rgb'] - o['ssim_rgb']:+.2e}  psnr_y diff {g['psnr_y'] - o['psnr_y']:+.2e}")
