"""Per-call host overhead of the C-ABI entry points on tiny frames (kernels ~ microseconds)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jpeg_dsp_studio_b200 as J
eng = J.Engine(0)
img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (64, 64, 3), dtype=np.uint8)).cuda()
rec = torch.empty_like(img)
def t(fn, n=300):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e6
print("roundtrip fast (device, no coeffs): %.1f us" % t(lambda: eng.roundtrip(img, 50, "4:2:0", False, precision="fast", recon_out=rec)))
print("roundtrip exact (device): %.1f us" % t(lambda: eng.roundtrip(img, 50, "4:2:0", False, precision="exact", recon_out=rec)))
print("sweep 13 pts fast (device): %.1f us" % t(lambda: eng.sweep(img, list(range(1, 14)), "4:2:0", False, precision="fast")))
fr = img[None].repeat(8, 1, 1, 1).contiguous(); out = torch.empty_like(fr)
print("batch 8 fast (device): %.1f us" % t(lambda: eng.roundtrip_batch(fr, 50, "4:2:0", False, precision="fast", recon_out=out)))
import ctypes as C
from jpeg_dsp_studio_b200 import _native as N
lib = N.load()
p = N.JdsParams(64, 64, 50, 2, 0, 1, 1 | 32 | 64, 0); m = N.JdsMetrics()
print("raw jds_roundtrip (ctypes only): %.1f us" % t(lambda: lib.jds_roundtrip(eng._ctx, C.byref(p), C.c_void_p(img.data_ptr()), 1, C.c_void_p(rec.data_ptr()), None, None, None, 1, C.byref(m))))
