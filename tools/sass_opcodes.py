"""SASS opcode histogram per kernel of libjds.so (evidence of what the kernels are made of):
    python tools/sass_opcodes.py [lib] > profiles/rN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "jpeg_dsp_studio_b200", "libjds.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
EVID = ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "IDP", "DFMA", "DADD", "DMUL", "IMMA",
        "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM")
print("# SASS opcode histogram per kernel (cuobjdump -sass jpeg_dsp_studio_b200/libjds.so)")
print("# Blackwell-native evidence: UTMALDG = cp.async.bulk.tensor (tensor-map TMA), UBLKCP = cp.async.bulk, SYNCS = mbarrier,")
print("# FFMA2/FADD2/FMUL2 = packed f32x2 (sm_100), IDP = dp4a")
cur, ops = None, collections.Counter()


def flush():
    if cur and sum(ops.values()):
        print(f"\n## {cur}\n   instructions: {sum(ops.values())}")
        print("   " + ", ".join(f"{k}:{v}" for k, v in ops.most_common(16)))
        ev = ", ".join(f"{k}={ops[k]}" for k in EVID if ops.get(k))
        print(f"   evidence: {ev}")


for ln in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", ln)
    if m:
        flush()
        cur, ops = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m:
        ops[m.group(1)] += 1
flush()
