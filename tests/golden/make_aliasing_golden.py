#!/usr/bin/env python
"""Generate tests/golden/aliasing.json by running the UNMODIFIED reference worker
(gui/dialogs/aliasing_demo_dialog.py::AliasingDemoWorker, imported through
oracle.reference_shim.load_aliasing_demo - PySide6 names stubbed, skimage stand-in).

Build container only (needs /root/reference and OpenCV):  python tests/golden/make_aliasing_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_shim  # noqa: E402
from tests import cases as C  # noqa: E402


def main():
    import cv2
    import numpy
    import scipy
    R = reference_shim.load_aliasing_demo()
    out = {"versions": {"numpy": numpy.__version__, "scipy": scipy.__version__, "cv2": cv2.__version__},
           "cases": []}
    for case in C.ALIASING_CASES:
        img = case.image()
        res = R.run_worker(img, case.quality)
        rec = {"name": case.name, "shape": list(img.shape), "quality": case.quality,
               "input_sha256": C.sha(img)}
        for k in ("recon_no_pf", "recon_pf", "diff_no_pf", "diff_pf"):
            rec[k + "_sha256"] = C.sha(res[k])
        for k in ("metrics_no_pf", "metrics_pf"):
            rec[k] = {n: (repr(v) if v != v or v in (float("inf"), float("-inf")) else float(v))
                      for n, v in res[k].items()}
        out["cases"].append(rec)
        print(case.name, rec["metrics_pf"])
    with open(os.path.join(HERE, "aliasing.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
