#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total ms, share."""
import collections
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Metric Name')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'^void ', '', r[ki])
        name = re.sub(r'\(.*', '', name)[:100]
        v = float(r[vi].replace(',', ''))
        ms = {'ns': v / 1e6, 'nsecond': v / 1e6, 'us': v / 1e3, 'usecond': v / 1e3, 'ms': v, 'msecond': v}.get(r[ui], v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print(f"{sum(a[0] for a in agg.values())} launches, {tot:.2f} ms total (serialised, cold-cache: compare shares)")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ms:10.3f} ms {100 * ms / tot:5.1f}%  x{n:4d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
