"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per-kernel launch count, total and mean device time and share of the step.

  python tools/summarize_ncu.py gpurun_out/launches.csv [--skip N] [--frames F] > profiles/rNN_launches.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    return name.strip()


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    frames = int(sys.argv[sys.argv.index("--frames") + 1]) if "--frames" in sys.argv else 1
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((short(r["Kernel Name"]), ns, r.get("Grid Size", ""), r.get("Block Size", "")))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, ns, grid, blk in rows:
        a = agg.setdefault(name, [0, 0.0, grid, blk])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"launches: {len(rows)} ({len(rows) / frames:.1f} per frame over {frames} frames); "
          f"sum of kernel time {total / 1e3 / frames:.1f} us per frame (cold-cache, serialised under ncu)\n")
    print("| kernel | launches/frame | us/frame | mean us | share | example grid x block |")
    print("|---|---:|---:|---:|---:|---|")
    for name, (n, ns, grid, blk) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n / frames:.1f} | {ns / 1e3 / frames:.1f} | {ns / 1e3 / n:.2f} | "
              f"{100 * ns / total:.1f}% | {grid} x {blk} |")


if __name__ == "__main__":
    main()
