"""Per-stage kernel times of one mode on 16 x 4K frames (A/B runs):
    [JDS_LIB=...] python tools/mode_time.py [fast|exact] [mode] [prefilter]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

os.environ.setdefault("JDS_SCRATCH_MB", "8192")
import jpeg_dsp_studio_b200 as J

prec = sys.argv[1] if len(sys.argv) > 1 else "fast"
mode = sys.argv[2] if len(sys.argv) > 2 else "4:2:0"
pf = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
frames = np.stack([np.random.default_rng(4000 + k).integers(0, 256, (2160, 3840, 3), dtype=np.uint8)
                   for k in range(16)])
d = torch.from_numpy(frames).cuda()
eng = J.Engine(0)
for _ in range(3):
    eng.roundtrip_batch(d, 50, mode, pf, precision=prec)
eng.stage_times(reset=True)
n = 10 if prec == "fast" else 4
for _ in range(n):
    outs = eng.roundtrip_batch(d, 50, mode, pf, precision=prec)
st = eng.stage_times()
tot = sum(v["ms"] for v in st.values()) / n
print(f"{os.path.basename(os.environ.get('JDS_LIB', 'libjds.so')):20s} {prec} {mode} pf={int(pf)}: "
      + " ".join(f"{k} {v['ms'] / n:.4f}" for k, v in st.items())
      + f" | total {tot:.4f} ms/16x4K = {16 * 3840 * 2160 / tot / 1e6:.1f} Gpx/s"
      + f" | psnr_y {outs[0].scalars['psnr_y']:.6f} ssim_y {outs[0].scalars['ssim_y']:.9f}")
