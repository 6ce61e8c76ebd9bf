"""Small end-to-end run for compute-sanitizer: every kernel family once on tiny frames."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jpeg_dsp_studio_b200 as J

rng = np.random.default_rng(0)
eng = J.get_engine(0)
for shape, mode, pf in [((48, 64), "4:2:0", False), ((40, 48), "4:2:2", True), ((32, 48), "4:4:4", False),
                        ((50, 70), "4:2:0", True), ((144, 272), "4:2:0", False)]:
    img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    for prec in ("exact", "fast"):
        o = eng.roundtrip(img, 50, mode, pf, precision=prec, want_coeffs=True,
                          want_error_maps=(prec == "exact"), want_hist=(prec == "exact"))
        print(shape, mode, pf, prec, round(o.scalars["psnr_y"], 3), round(o.scalars["ssim_y"], 5))
frames = np.stack([rng.integers(0, 256, (64, 96, 3), dtype=np.uint8) for _ in range(5)])
print(len(eng.roundtrip_batch(frames, 30, "4:2:2", False, precision="fast")))
print(len(eng.sweep(frames[0], [5, 50, 95], "4:2:0", False, precision="fast", want_recon=True)))
res, inter = J.compress_reconstruct(frames[1], J.CompressionParams(quality=70), (1, 2))
print(res.psnr_y, inter.selected_block_dct[0, 0])
from jpeg_dsp_studio_b200 import engines as E
print(E.dct2(np.ones((8, 8)))[0, 0])
