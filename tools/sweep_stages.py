"""Per-stage kernel time of a 13-point and a 100-point sweep (CUDA events around every kernel):
where the non-scaling part of an 8-GPU rank's share comes from."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
os.environ.setdefault("JDS_SCRATCH_MB", "8192")
import jpeg_dsp_studio_b200 as J
eng = J.Engine(0)
img = torch.from_numpy(np.random.default_rng(4).integers(0, 256, (2160, 3840, 3), dtype=np.uint8)).cuda()
q100 = list(range(1, 101))
for qs in (q100, q100[::8][:13], q100[::8][:12]):
    for _ in range(3):
        eng.sweep(img, qs, "4:2:0", False, precision="fast")
    eng.set_stage_timing(True)
    eng.stage_times(reset=True)
    l0 = eng.launch_count()
    reps = 10
    for _ in range(reps):
        eng.sweep(img, qs, "4:2:0", False, precision="fast")
    st = eng.stage_times(reset=True)
    eng.set_stage_timing(False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.sweep(img, qs, "4:2:0", False, precision="fast")
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    n = len(qs)
    print(f"{n:3d} points: wall {wall:.4f} ms ({wall / n:.4f}/pt); launches/sweep {(eng.launch_count() - l0) / (2 * reps):.1f}; " +
          "; ".join(f"{k} {v['ms'] / reps:.4f} ms ({v['ms'] / reps / n:.4f}/pt, {v['launches'] / reps:.0f} launches)" for k, v in st.items()))
