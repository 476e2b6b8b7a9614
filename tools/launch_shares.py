#!/usr/bin/env python
"""Per-kernel share of GPU time from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys


def main(path):
    tot = collections.Counter()
    cnt = collections.Counter()
    for row in csv.reader(open(path, errors="replace")):
        if len(row) > 14 and row[12] == "gpu__time_duration.sum":
            name = re.sub(r"\(.*", "", row[4]).replace("vrdx::", "").replace("void ", "")
            tot[name[:70]] += float(row[14])
            cnt[name[:70]] += 1
    whole = sum(tot.values())
    print("share of GPU time per kernel in the launch list (cold-cache, serialised under ncu: compare shares, not absolutes)")
    for name, ns in tot.most_common():
        if ns / whole >= 0.002:
            print(f"{100 * ns / whole:6.2f}%  {ns / 1e6:9.3f} ms  {cnt[name]:4d} launches  {name}")


if __name__ == "__main__":
    main(sys.argv[1])
