#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`)."""
import collections
import csv
import json
import re
import sys


def main():
    path, passes = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        rows.append((name, ms))
    tot = sum(ms for _, ms in rows)
    by = collections.defaultdict(lambda: [0, 0.0])
    for n, ms in rows:
        by[n][0] += 1
        by[n][1] += ms
    out = {"launches": len(rows), "kernel_ms_total": tot, "passes": passes, "launches_per_pass": len(rows) / passes, "kernel_ms_per_pass": tot / passes,
           "kernels": [{"kernel": n, "launches_per_pass": c / passes, "ms_per_pass": ms / passes, "share": ms / tot} for n, (c, ms) in
                       sorted(by.items(), key=lambda kv: -kv[1][1])[:25]]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
