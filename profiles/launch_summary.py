"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:

    python profiles/launch_summary.py gpurun_out/launches.csv > profiles/rN/launches_....summary.txt
"""
import collections
import csv
import re
import sys


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val / 1e3 if unit in ("ns", "nsecond") else val * {"us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, us))
    agg = collections.OrderedDict()
    for name, us in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += us
    total = sum(v[1] for v in agg.values()) or 1.0
    print("%-70s %8s %12s %8s %10s" % ("kernel", "launches", "sum us", "share", "mean us"))
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %8d %12.1f %7.1f%% %10.2f" % (name[:70], cnt, us, 100 * us / total, us / cnt))
    print("%-70s %8d %12.1f" % ("TOTAL", len(rows), total))


if __name__ == "__main__":
    main(sys.argv[1])
