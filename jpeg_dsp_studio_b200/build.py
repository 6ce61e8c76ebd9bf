"""Build libjds.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Used by ``__graft_entry__.build()`` and runnable by hand:
    python jpeg_dsp_studio_b200/build.py [--force] [--verbose]
nvcc cross-compiles without a GPU; the resulting ``jpeg_dsp_studio_b200/libjds.so``
is git-ignored but travels to the GPU box with the repo snapshot.
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjds.so")
SOURCES = ["jds_api.cu", "jds_kernels.cu", "jds_ssim.cu", "jds_fused.cu", "jds_fused_exact.cu", "jds_preview.cu", "jds_ops.cu", "jds_alias.cu", "jds_entropy.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "jds.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    """Compile csrc/*.cu into libjds.so; returns the library path."""
    if not force and not _stale():
        return LIB
    cmd = [_nvcc()] + ARCH_FLAGS + NVCC_FLAGS
    if verbose:
        cmd += ["-Xptxas", "-v"]
    tmp = LIB + ".tmp"
    cmd += ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libjds.so")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
