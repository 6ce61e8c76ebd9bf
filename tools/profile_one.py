"""One short run of the hot path for ncu: `python tools/profile_one.py [fast|exact] [frames]`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
os.environ.setdefault("JDS_SCRATCH_MB", "8192")      # like bench.py: the whole batch in one launch sequence
import jpeg_dsp_studio_b200 as J

precision = sys.argv[1] if len(sys.argv) > 1 else "fast"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "4:2:0"
frames = np.stack([np.random.default_rng(4000 + k).integers(0, 256, (2160, 3840, 3), dtype=np.uint8)
                   for k in range(n)])
d = torch.from_numpy(frames).cuda()
eng = J.Engine(0)
for _ in range(2):
    outs = eng.roundtrip_batch(d, 50, mode, False, precision=precision)
print(outs[0].scalars["psnr_y"], outs[0].scalars["ssim_y"], eng.stage_times())
