#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv, collections, sys, json
def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    return list(csv.DictReader(lines))
def main():
    rows = load(sys.argv[1])
    per_pass = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    d = collections.OrderedDict()
    for row in rows:
        k = row['Kernel Name']
        k = k.split('(')[0][:70]
        v = float(row['Metric Value'].replace(',', ''))
        if row['Metric Unit'] in ('us', 'usecond'): v *= 1e3
        elif row['Metric Unit'] in ('ms', 'msecond'): v *= 1e6
        d.setdefault(k, [0, 0.0]); d[k][0] += 1; d[k][1] += v
    tot = sum(t for _, t in d.values())
    print(f"launches {len(rows)}  total {tot/1e6:.3f} ms (all passes)")
    for k, (c, t) in sorted(d.items(), key=lambda x: -x[1][1]):
        print(f"{k:70s} {c:5d} {t/1e6:9.3f} ms  {100*t/tot:5.1f}%  avg {t/c/1e3:9.1f} us")
main()
