import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
os.environ.setdefault("JDS_SCRATCH_MB", "8192")
import jpeg_dsp_studio_b200 as J
n = int(sys.argv[1])
frames = np.stack([np.random.default_rng(4000 + k).integers(0, 256, (2160, 3840, 3), dtype=np.uint8) for k in range(n)])
d = torch.from_numpy(frames).cuda()
eng = J.Engine(0)
for _ in range(3): eng.roundtrip_batch(d, 50, "4:2:0", False, precision="fast")
eng.stage_times(reset=True)
for _ in range(10): eng.roundtrip_batch(d, 50, "4:2:0", False, precision="fast")
st = eng.stage_times()
print(f"units {n} segs {os.environ.get('JDS_SSIM_SEGS','auto')}: ssim {st['ssim']['ms']/10:.4f} ms = {st['ssim']['ms']/10/n:.5f} per frame")
