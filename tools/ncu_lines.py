#!/usr/bin/env python
"""Per-CUDA-source-line instruction counts / stall samples from an ncu report.

    python tools/ncu_lines.py report.ncu-rep libjds.so kernel_substring [top_n]

ncu's CSV source page is SASS level only; this joins it (by instruction order) with
`nvdisasm -g` line info of the same cubin and aggregates by source line.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def disasm_lines(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True,
                   stdout=subprocess.DEVNULL)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin") or f.count("-") > 0:
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True,
                             text=True).stdout
        if kernel not in txt:
            continue
        cur, infunc = None, False
        for ln in txt.splitlines():
            if ln.startswith("//----") and ".text." in ln:
                infunc = kernel in ln
                continue
            if not infunc:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out.append((cur, m.group(2).strip()))
        if out:
            break
    return out


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis_key = sys.argv[5] if len(sys.argv) > 5 else kernel      # mangled-name substring for nvdisasm
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kernel}"],
                            capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(csvtxt)))
    # the first kernel block only: header at row 1, a new "Kernel Name" row starts the next
    end = next((i for i, r in enumerate(rows[2:], 2) if r and r[0] == "Kernel Name"), len(rows))
    rows = rows[:end]
    hdr = rows[1]
    iS, iI, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    sass = [(r[iS].strip(), int(r[iI] or 0), int(r[iW] or 0)) for r in rows[2:] if len(r) > max(iW, iI) and r[iI].isdigit()]
    dis = disasm_lines(lib, dis_key)
    if len(dis) != len(sass):
        print(f"warning: {len(dis)} disassembled vs {len(sass)} profiled instructions (library rebuilt?)")
    per = collections.defaultdict(lambda: [0, 0])
    tot = sum(s[1] for s in sass)
    tots = sum(s[2] for s in sass)
    for (loc, _), (_, n, w) in zip(dis, sass):
        per[loc][0] += n
        per[loc][1] += w
    srcs = {}
    print(f"total warp-instructions {tot}, stall samples {tots}")
    for loc, (n, w) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if loc:
            path = None
            for root in ("jpeg_dsp_studio_b200/csrc", "."):
                p = os.path.join(root, loc[0])
                if os.path.exists(p):
                    path = p
                    break
            if path:
                srcs.setdefault(path, open(path).read().splitlines())
                if loc[1] - 1 < len(srcs[path]):
                    text = srcs[path][loc[1] - 1].strip()[:90]
        print(f"{100 * n / tot:5.1f}% instr {100 * w / max(tots, 1):5.1f}% stall  {loc}  {text}")


if __name__ == "__main__":
    main()
