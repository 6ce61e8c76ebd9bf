#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv):
    python tools/launch_summary.py gpurun_out/launches_r1.csv "title" > profiles/rN_launches_summary.txt"""
import collections
import csv
import sys


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        name = r[ik].split("(")[0][:60]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {title}")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"{'kernel':62s} {'launches':>8s} {'total us':>10s} {'share':>7s} {'avg us':>9s}")
    for k, v in tot.most_common():
        print(f"{k:62s} {cnt[k]:8d} {v:10.1f} {100 * v / total:6.1f}% {v / cnt[k]:9.1f}")


if __name__ == "__main__":
    main()
