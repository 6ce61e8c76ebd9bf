"""DRAM traffic of the hot path per kernel and for the whole step, from a LIVE one-pass ncu run:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --cache-control none --clock-control none --csv --log-file gpurun_out/traffic.csv \
        python tools/profile_one.py fast 16
    python tools/traffic.py gpurun_out/traffic.csv 16 [label] > profiles/r2_traffic.json

Two metrics fit one pass, so nothing is replayed and `--cache-control none` leaves L2 exactly as
the preceding kernels left it: the figures are the traffic of the kernels running back to back
(the LAST batch call of the script is taken: chroma, luma, SSIM kernels of `frames` 4K frames).
"""
import csv
import json
import sys


def main():
    path, frames = sys.argv[1], int(sys.argv[2])
    label = sys.argv[3] if len(sys.argv) > 3 else "default"
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ik, im, iv, iu, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    launches = {}
    for r in rows[hdr + 1:]:
        if len(r) <= iv:
            continue
        d = launches.setdefault(int(r[iid]), {"kernel": r[ik]})
        v = float(r[iv].replace(",", ""))
        unit = r[iu].lower()
        if r[im].startswith("dram__bytes"):
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        d[r[im]] = v
    ids = sorted(launches)
    # tools/profile_one.py makes two identical batch calls: the second half of the launches is the
    # second call (caches warm, scratch allocated)
    last = [launches[i] for i in ids[len(ids) // 2:]]
    px = frames * 3840 * 2160
    out = {"source": f"ncu one-pass live run (--cache-control none, no replay) of tools/profile_one.py fast {frames}: "
                     f"the kernels of one {frames} x 4K batch call back to back, mode {label}",
           "pixels": px, "kernels": []}
    tot = 0.0
    per = {}
    for d in last:
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        tot += b
        name = d["kernel"].split("(")[0].replace("void ", "").split("<")[0].replace("jds::", "")
        k = per.setdefault(name, {"kernel": name, "launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0,
                                  "gpu_time_us": 0.0})
        k["launches"] += 1
        k["dram_bytes_read"] += d.get("dram__bytes_read.sum", 0.0)
        k["dram_bytes_write"] += d.get("dram__bytes_write.sum", 0.0)
        k["gpu_time_us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3      # ns -> us
    for name, k in per.items():
        k["dram_bytes_per_pixel"] = round((k["dram_bytes_read"] + k["dram_bytes_write"]) / px, 3)
        k["gpu_time_us"] = round(k["gpu_time_us"], 1)
        out["kernels"].append(k)
        out[name] = {"dram_bytes_per_pixel": k["dram_bytes_per_pixel"], "pixels_in_launch": px // max(k["launches"], 1)}
    out["whole_path"] = {"dram_bytes_per_pixel": round(tot / px, 3), "algorithmic_bytes_per_pixel": 6.0,
                         "source": out["source"]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
