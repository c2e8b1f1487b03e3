#!/usr/bin/env python
"""Per-kernel launch count / mean / min duration from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`).

    python tools/launch_summary.py gpurun_out/X_launches.csv
"""
import collections
import csv
import sys


def summarize(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    else:
        raise SystemExit("no ncu CSV header in " + path)
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[start:]:
        if len(r) > mv:
            try:
                d[r[kn].split("(")[0]].append(float(r[mv].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:28s} n={len(v):4d} mean={sum(v) / len(v) / 1e3:9.1f} us  min={min(v) / 1e3:9.1f} us  share={100 * sum(v) / tot:5.1f} %")


if __name__ == "__main__":
    summarize(sys.argv[1])
