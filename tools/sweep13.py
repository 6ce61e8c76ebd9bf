"""Where does an 8-GPU rank's share of a 100-point sweep (13 points) spend its time?  One GPU,
13 points: CUDA-event span of the device-resident sweep (kernels + record packing + D2H) against
the per-point cost taken from a 100-point sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
os.environ.setdefault("JDS_SCRATCH_MB", "8192")
import jpeg_dsp_studio_b200 as J
from jpeg_dsp_studio_b200 import distributed as D
dev = torch.device("cuda", 0)
eng = J.Engine(0)
stream = torch.cuda.current_stream(dev)
eng.use_stream(stream.cuda_stream)
img = torch.from_numpy(np.random.default_rng(4).integers(0, 256, (2160, 3840, 3), dtype=np.uint8)).to(dev)
def span(qs, reps=20):
    for _ in range(3):
        D.sweep_sharded_begin(eng, img, qs, "4:2:0", False, precision="fast", device=dev).result()
    tot = 0.0
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        h = D.sweep_sharded_begin(eng, img, qs, "4:2:0", False, precision="fast", device=dev)
        b.record(stream)
        h.result()
        tot += a.elapsed_time(b)
    return tot / reps
q100 = list(range(1, 101))
t100 = span(q100)
for n in (12, 13, 25, 50):
    qs = q100[::8][:n] if n <= 13 else q100[:n]
    t = span(qs)
    print(f"{n:3d} points: {t:.4f} ms device span; per point {t / n:.4f}; 100-point sweep {t100:.4f} ms -> "
          f"{t100 / 100:.4f} per point; fixed part of the {n}-point sweep ~ {t - n * t100 / 100:.4f} ms")
