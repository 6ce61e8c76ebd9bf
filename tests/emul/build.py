"""Build tests/emul/libjds_emul.so (TEST INFRASTRUCTURE): the product's device
functions compiled for the host so the CPU suite can check them against the oracle."""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libjds_emul.so")
SRC = os.path.join(HERE, "emul.cpp")
CSRC = os.path.join(ROOT, "jpeg_dsp_studio_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    # no -march=native / -mfma: fp64 ops must stay individually rounded
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
           "-o", LIB + ".tmp", SRC]
    subprocess.run(cmd, check=True)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
