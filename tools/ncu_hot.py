#!/usr/bin/env python
"""Summarise the hottest SASS instructions of an ncu report's source page.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [top=40]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        n = int(r[col['# Samples']])
    except ValueError:
        continue
    data.append((n, r))
total = sum(n for n, _ in data)
print('total samples', total)
agg = {h: 0 for h in stall_cols}
for n, r in data:
    for h in stall_cols:
        try:
            agg[h] += int(r[col[h]])
        except ValueError:
            pass
print('stall totals:', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
for i in sorted(order):
    n, r = data[i]
    st = {h[6:]: int(r[col[h]]) for h in stall_cols if r[col[h]] not in ('', '0')}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f'{i:5d} {n:6d} {100 * n / total:5.1f}%  exec {r[col["Instructions Executed"]]:>9}  {r[col["Source"]][:90]:90s} {st}')
