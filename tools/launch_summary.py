#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel.  usage: launch_summary.py launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    k = r['Kernel Name'][:60]
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += float(r['Metric Value']) / 1000
tot = sum(v[1] for v in agg.values())
print(f'total us {tot:.1f}  ({len(rows)} launches)')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{v[1]:9.1f} us {v[0]:4d} launches {v[1] / v[0]:8.2f} us/launch {100 * v[1] / tot:5.1f}%  {k}')
