"""Stage the UNMODIFIED reference for the CPU arm of the benchmark (TEST INFRASTRUCTURE).

    python oracle/make_ref.py            # /root/reference -> oracle/_ref/reference

`/root/reference` exists only in the build container; the GPU box gets the repository
snapshot.  `oracle/_ref/` is git-ignored (no reference source enters the history) but not
gpurun-ignored, so the staged files travel with the snapshot and `bench.py --impl reference`
can time the reference AS IS on the box's host cores (`cpu_baseline.kind == "reference"`,
SURVEY.md §8d "baseline of record").  Only the files of the hot path are staged - the three
packages `engines/pipeline.py:1-11` imports - byte for byte, together with a manifest of their
SHA-256 digests.  scikit-image, which `utils/metrics.py:5` imports, is not installed anywhere
here: `oracle/reference_shim.py` registers `oracle/skimage_standin.py` before importing, on
the box exactly as in the build container.

`__graft_entry__.build()` calls `stage()` whenever `/root/reference` is present.
"""

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("JDS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
PACKAGES = ("engines", "models", "utils")


def staged_root():
    """Path of the staged reference, or None."""
    return DST if os.path.isfile(os.path.join(DST, "engines", "pipeline.py")) else None


def stage(verbose=False):
    """Copy the hot path's packages; returns the destination or None when there is no source."""
    if not os.path.isfile(os.path.join(SRC, "engines", "pipeline.py")):
        return staged_root()
    manifest = {}
    for pkg in PACKAGES:
        os.makedirs(os.path.join(DST, pkg), exist_ok=True)
        for name in sorted(os.listdir(os.path.join(SRC, pkg))):
            if not name.endswith(".py"):
                continue
            a, b = os.path.join(SRC, pkg, name), os.path.join(DST, pkg, name)
            shutil.copyfile(a, b)
            manifest[f"{pkg}/{name}"] = hashlib.sha256(open(b, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"[oracle/_ref] staged {len(manifest)} reference files into {DST}")
    return DST


if __name__ == "__main__":
    stage(verbose=True)
