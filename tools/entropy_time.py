"""Device-side entropy coder (jds_entropy_encode): ms per 4K frame, device-resident coefficients,
against the size-only kernel and the raw size of the coefficient array.
    python tools/entropy_time.py"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import jpeg_dsp_studio_b200 as J
from jpeg_dsp_studio_b200 import _native as N
from tests import cases as CS

eng = J.Engine(0)
rows = []
for name, img in (("photo_tiled", CS.photo_tiled(2160, 3840)),
                  ("uniform_random", np.random.default_rng(3).integers(0, 256, (2160, 3840, 3), dtype=np.uint8))):
    for q in (20, 50, 90):
        o = eng.roundtrip(torch.from_numpy(img).cuda(), q, "4:2:0", False, precision="fast", want_coeffs=True)
        c = o.coeffs
        nb, bits = (C.c_uint64 * 3)(), (C.c_uint64 * 3)()
        outbuf = torch.empty(2 * c.numel(), dtype=torch.uint8, device="cuda")
        def enc():
            N.check(eng._lib.jds_entropy_encode(eng._ctx, C.c_void_p(c.data_ptr()), N.JDS_DEVICE, 2160, 3840,
                                                N.JDS_SUB_420, C.c_void_p(outbuf.data_ptr()), N.JDS_DEVICE,
                                                outbuf.numel(), nb, bits))
        def size_only():
            eng.entropy_bits(c, 2160, 3840, "4:2:0")
        res = {}
        for label, fn in (("encode_ms", enc), ("size_only_ms", size_only)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            res[label] = round((time.perf_counter() - t0) * 50, 4)
        total = int(sum(nb))
        rows.append({"content": name, "quality": q, **res, "scan_bytes": total,
                     "bpp": round(8 * total / (2160 * 3840), 4), "coeff_bytes": 2 * c.numel(),
                     "coeff_GBps": round(2 * c.numel() / res["encode_ms"] / 1e6, 1),
                     "Mpx_per_s": round(2160 * 3840 / res["encode_ms"] / 1e3, 1)})
print(json.dumps({"frame": "2160x3840 4:2:0", "note": "wall clock per call incl. the two size read-backs the call synchronises on", "rows": rows}))
