"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<round>_launches.md

Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live
CUDA-event shares (roofline.kernels[*].share_of_step), not absolutes (B200_PROFILING.md)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", row["Kernel Name"]))
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
        a = agg.setdefault((name, row["Grid Size"], row["Block Size"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"source: {path}  ({sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms summed)\n")
    print("| kernel | grid | block | launches | total us | avg us | share |")
    print("|---|---|---|---:|---:|---:|---:|")
    for (k, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {g} | {b} | {n} | {t:.1f} | {t / n:.1f} | {t / tot:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1])
