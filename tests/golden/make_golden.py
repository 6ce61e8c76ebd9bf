#!/usr/bin/env python
"""Generate tests/golden/*.json by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # all cases (4K ones take ~25 s each)
    python tests/golden/make_golden.py --quick    # skip cases marked big

The reference's ``compress_reconstruct`` (engines/pipeline.py:17-167) is imported
through ``oracle.reference_shim`` (skimage stand-in injected, nothing else changed)
and executed on every case in ``tests/cases.py``.  For each case the file records
the SHA-256 of the input, of ``all_quantized_coeffs`` (int16), of
``reconstructed_image`` (uint8), of both error maps (fp64 bytes), the 50-bin
histogram, the six selected-block arrays and every scalar of CompressionResult.
Library versions are recorded because the reference's arithmetic is whatever
numpy / scipy / cv2 are installed (SURVEY.md §7 'library drift').
"""

import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_shim  # noqa: E402
from tests import cases as C  # noqa: E402


def run_case(R, case):
    img = case.image()
    params = R.CompressionParams(quality=case.quality, subsampling_mode=case.mode,
                                 use_prefilter=case.prefilter)
    t0 = time.perf_counter()
    res, inter = R.compress_reconstruct(img, params, case.sel)
    dt = time.perf_counter() - t0
    rec = {
        "name": case.name, "shape": list(img.shape), "quality": case.quality,
        "mode": case.mode, "prefilter": case.prefilter, "sel": list(case.sel),
        "input_sha256": C.sha(img),
        "coeffs_sha256": C.sha(inter.all_quantized_coeffs),
        "coeffs_dtype": str(inter.all_quantized_coeffs.dtype),
        "coeffs_len": int(inter.all_quantized_coeffs.size),
        "recon_sha256": C.sha(res.reconstructed_image),
        "error_map_y_sha256": C.sha(inter.error_map_y),
        "error_map_rgb_sha256": C.sha(inter.error_map_rgb),
        "histogram": [int(v) for v in inter.quantized_histogram],
        "psnr_y": res.psnr_y, "ssim_y": res.ssim_y,
        "psnr_rgb": res.psnr_rgb, "ssim_rgb": res.ssim_rgb,
        "bpp": res.bpp, "compression_ratio": res.compression_ratio,
        "nonzero_coeffs": res.nonzero_coeffs, "total_coeffs": res.total_coeffs,
        "bitrate_label": res.bitrate_label,
        "reference_seconds": round(dt, 3),
    }
    if inter.selected_block_dct is None:
        rec["selected"] = None
    else:
        rec["selected"] = {
            k: C.sha(getattr(inter, "selected_block_" + k))
            for k in ("original", "shifted", "dct", "quantized", "dequantized",
                      "reconstructed")}
        rec["selected_quantized"] = inter.selected_block_quantized.astype(int).ravel().tolist()
    return rec


def jsonable(x):
    if isinstance(x, float) and (x != x or x in (float("inf"), float("-inf"))):
        return repr(x)
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    R = reference_shim.load()
    import cv2
    import scipy
    meta = {
        "generator": "tests/golden/make_golden.py",
        "reference_root": reference_shim.REFERENCE_ROOT,
        "numpy": np.__version__, "scipy": scipy.__version__, "cv2": cv2.__version__,
        "cv2_ipp": bool(cv2.ipp.useIPP()), "cv2_ipp_version": cv2.ipp.getIppVersion(),
        "ssim_source": "oracle/skimage_standin.py (scikit-image not installed)",
    }
    path = os.path.join(HERE, "cases.json")
    old = {}
    if os.path.exists(path):
        old = {r["name"]: r for r in json.load(open(path))["cases"]}
    out = []
    for case in C.CASES:
        if args.only and case.name != args.only:
            if case.name in old:
                out.append(old[case.name])
            continue
        if args.quick and case.big and case.name in old:
            out.append(old[case.name])
            continue
        rec = run_case(R, case)
        rec = {k: jsonable(v) for k, v in rec.items()}
        print(f"{case.name:32s} {rec['reference_seconds']:7.2f}s psnr_y={rec['psnr_y']}", flush=True)
        out.append(rec)
    json.dump({"meta": meta, "cases": out}, open(path, "w"), indent=1)

    if not args.only:
        name, make = C.SWEEP_IMAGE
        img = make()
        sweep = {"image": name, "input_sha256": C.sha(img), "points": []}
        for mode, pf in C.SWEEP_COMBOS:
            for q in C.SWEEP_QUALITIES:
                res, inter = R.compress_reconstruct(
                    img, R.CompressionParams(quality=q, subsampling_mode=mode, use_prefilter=pf))
                sweep["points"].append({
                    "mode": mode, "prefilter": pf, "quality": q,
                    "coeffs_sha256": C.sha(inter.all_quantized_coeffs),
                    "recon_sha256": C.sha(res.reconstructed_image),
                    "psnr_y": jsonable(res.psnr_y), "ssim_y": res.ssim_y,
                    "psnr_rgb": jsonable(res.psnr_rgb), "ssim_rgb": res.ssim_rgb,
                    "bpp": res.bpp, "compression_ratio": res.compression_ratio,
                    "nonzero_coeffs": res.nonzero_coeffs})
            print("sweep", mode, pf, "done", flush=True)
        json.dump({"meta": meta, "sweep": sweep},
                  open(os.path.join(HERE, "sweep.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
