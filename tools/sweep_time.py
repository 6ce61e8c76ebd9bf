"""100-point sweep of one 4K frame on one GPU: ms per sweep and a checksum of the table
    [JDS_NO_HOIST=1] python tools/sweep_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
os.environ.setdefault("JDS_SCRATCH_MB", "8192")
import jpeg_dsp_studio_b200 as J
img = np.random.default_rng(4).integers(0, 256, (2160, 3840, 3), dtype=np.uint8)
d = torch.from_numpy(img).cuda()
eng = J.Engine(0)
qs = list(range(1, 101))
for pf in (False, True):
    for _ in range(3):
        outs = eng.sweep(d, qs, "4:2:0", pf, precision="fast")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        outs = eng.sweep(d, qs, "4:2:0", pf, precision="fast")
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 100
    chk = sum(o.scalars["psnr_y"] + o.scalars["ssim_y"] + o.scalars["bpp"] for o in outs)
    print(f"hoist={'0' if os.environ.get('JDS_NO_HOIST') else '1'} pf={int(pf)}: {ms:.3f} ms per 100-point sweep, checksum {chk:.9f}")
