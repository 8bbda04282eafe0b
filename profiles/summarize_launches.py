"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py <launches.csv> [filter-substring]"""
import collections
import csv
import re
import sys


def main(path, flt=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        if flt and flt not in k:
            continue
        print(f"{sum(v):10.1f} us  n={len(v):4d}  avg={sum(v)/len(v):8.2f} us  {100*sum(v)/tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
