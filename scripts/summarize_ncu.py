"""Summarise an `ncu --csv --log-file` launch list (one gpu__time_duration.sum row per launch) into
kernel, launches, total_us, share -- the table committed under profiles/."""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("<unnamed>::", "")
    return re.sub(r"\(.*$", "", name)[:140]


def main(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            val = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit.startswith("u") else val * 1e3)
            rows.append((short(r["Kernel Name"]), us))
    agg = defaultdict(lambda: [0, 0.0])
    for k, us in rows:
        agg[k][0] += 1
        agg[k][1] += us
    total = sum(v[1] for v in agg.values()) or 1.0
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share"])
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, "%.1f" % us, "%.4f" % (us / total)])
    print("%d launches, %.1f us total, %d kernels -> %s" % (len(rows), total, len(agg), dst))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
