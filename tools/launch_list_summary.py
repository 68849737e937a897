#!/usr/bin/env python
"""Kernel shares from an ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...).
Usage: python tools/launch_list_summary.py X.csv [header comment]   -> table on stdout (per-launch times are cold-cache and
serialised: compare SHARES, not absolutes)."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = None
    for r in rd:
        if header is None:
            if "Kernel Name" in r and "Metric Value" in r:
                header = r
            continue
        if len(r) != len(header):
            continue
        d = dict(zip(header, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "ns")
        us = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*$", "", d["Kernel Name"]).replace("void ", "").strip()
        rows.append((name, us))
    tot = sum(u for _, u in rows) or 1.0
    agg = collections.OrderedDict()
    for name, us in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    if len(sys.argv) > 2:
        print("# " + " ".join(sys.argv[2:]))
    print("# %d launches, %.1f ms summed; per-launch times are cold-cache and serialised: compare SHARES, not absolutes" % (len(rows), tot * 1e-3))
    print("%-70s %9s %12s %7s" % ("kernel", "launches", "total_us", "share"))
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %9d %12.1f %6.1f%%" % (name[:70], cnt, us, 100.0 * us / tot))


if __name__ == "__main__":
    main()
