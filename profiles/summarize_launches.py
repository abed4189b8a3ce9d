"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel
launch count, total / mean device time and share of the captured region."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void\s+|genie::\(anonymous namespace\)::|genie::", "", name)
        rows.append((name, v * scale))
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, t in rows:
        agg[n][0] += 1
        agg[n][1] += t
    print(f"# {path}: {len(rows)} launches, {tot / 1e3:.2f} ms device time (cold-cache, serialised: compare shares)")
    print(f"{'kernel':60s} {'launches':>9s} {'total_ms':>10s} {'mean_us':>9s} {'share':>7s}")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:60]:60s} {c:9d} {t / 1e3:10.3f} {t / c:9.2f} {100 * t / tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
