"""SSIM-kernel time and result of the library named by JDS_LIB (A/B runs of tools/ssim_variants.sh):
16 x 4K random frames, fast mode; prints ms per launch of k_ssim_strip and the SSIM values."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

os.environ.setdefault("JDS_SCRATCH_MB", "8192")
import jpeg_dsp_studio_b200 as J

frames = np.stack([np.random.default_rng(4000 + k).integers(0, 256, (2160, 3840, 3), dtype=np.uint8)
                   for k in range(16)])
d = torch.from_numpy(frames).cuda()
eng = J.Engine(0)
for _ in range(3):
    eng.roundtrip_batch(d, 50, "4:2:0", False, precision="fast")
eng.stage_times(reset=True)
for _ in range(10):
    outs = eng.roundtrip_batch(d, 50, "4:2:0", False, precision="fast")
st = eng.stage_times()
# a dark flat-ish frame: worst case for fp32 window sums
dark = (np.random.default_rng(1).integers(0, 6, (1, 1080, 1920, 3), dtype=np.uint8))
o2 = eng.roundtrip_batch(torch.from_numpy(dark).cuda(), 50, "4:2:0", False, precision="fast")
ex = eng.roundtrip_batch(torch.from_numpy(dark).cuda(), 50, "4:2:0", False, precision="exact")
print(f"{os.path.basename(os.environ.get('JDS_LIB', 'libjds.so')):24s} ssim {st['ssim']['ms'] / st['ssim']['launches']:.4f} ms "
      f"ssim_y {outs[0].scalars['ssim_y']:.9f} ssim_rgb {outs[0].scalars['ssim_rgb']:.9f} "
      f"dark: fast-exact ssim_y {o2[0].scalars['ssim_y'] - ex[0].scalars['ssim_y']:+.2e}")
