"""Wall-clock latency of the drop-in call at several sizes (exact / fast, with / without intermediates)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import jpeg_dsp_studio_b200 as J
for (h, w) in [(512, 512), (1080, 1920), (2160, 3840)]:
    img = np.random.default_rng(1).integers(0, 256, (h, w, 3), dtype=np.uint8)
    for prec in ("exact", "fast"):
        for inter in (True, False):
            p = J.CompressionParams(quality=50, subsampling_mode="4:2:0" if h % 16 == 0 else "4:2:2")
            for _ in range(2):
                J.compress_reconstruct(img, p, precision=prec, intermediates=inter)
            t = []
            for _ in range(5):
                t0 = time.perf_counter()
                r, i = J.compress_reconstruct(img, p, precision=prec, intermediates=inter)
                t.append(time.perf_counter() - t0)
            print(f"{w}x{h} {prec:5s} intermediates={inter!s:5s}: {1e3*min(t):8.3f} ms wall, kernels {r.encode_time_ms:.3f} ms "
                  f"-> {h*w/min(t)/1e6:9.1f} Mpixel/s")
