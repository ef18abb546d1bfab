#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, mean time, share."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(list)
    for r in rows:
        agg[r["Kernel Name"].split("(")[0]].append(float(r["Metric Value"]))
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot/1e6:.3f} ms of kernel time (cold-cache, serialised: compare shares)")
    print("| kernel | launches | mean us | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {k} | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/tot:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1])
