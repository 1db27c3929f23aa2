#!/usr/bin/env python
"""Pretty-print the clock64 event logs that tools/kbench_mlp.py writes (MLP_TRACE=1): per role, tag@cycle(+delta)."""
import sys

for l in open(sys.argv[1]).read().splitlines():
    parts = l.split()
    if not parts or parts[0] not in ('mma', 'epi', 'epi1'):
        continue
    ev = [tuple(map(int, x.split('@'))) for x in parts[1:] if '@' in x and x.split('@')[1].isdigit()]
    prev, out = None, []
    for tag, c in ev:
        out.append(f'{tag}@{c}' + (f'(+{c - prev})' if prev is not None else ''))
        prev = c
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**12
    print(parts[0], ' '.join(o for o, (t, c) in zip(out, ev) if lo <= c <= hi))
    print()
