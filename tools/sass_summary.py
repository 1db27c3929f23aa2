"""Static evidence from the built library (no GPU needed): per kernel of librovitkan.so the register / shared-memory / spill
figures ptxas recorded (cuobjdump -res-usage) and the counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UTMAPF = TMA prefetch,
SYNCS = mbarrier, MUFU = special-function unit; STL / LDL = local-memory stores / loads: STL without LDL is the
argument buffer of the mbarrier-timeout `printf`; LDL > 0 means a few values spilled at the register cap are read back).

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rovitkan_b200 import _lib  # noqa: E402

MNEMONICS = ('UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UTCBAR', 'SYNCS', 'MUFU', 'HMMA', 'FFMA', 'RED', 'ATOM', 'STL', 'LDL')


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = re.sub(r'^void ', '', n)
        n = re.sub(r'\((?:[^()]|\([^()]*\))*\)$', '', n)           # drop the parameter list
        n = n.replace('(anonymous namespace)::', '')
        short.append(n)
    return dict(zip(names, short))


def main():
    lib = _lib.LIB_PATH
    res = subprocess.run(['cuobjdump', '-res-usage', lib], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r'Function (\S+):\s*\n\s*(REG:\d+[^\n]*)', res):
        usage[m.group(1)] = dict(kv.split(':') for kv in m.group(2).split())
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    counts, total, cur, arch = {}, collections.Counter(), None, set()
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r'\s*arch = (\S+)', line)
        if m:
            arch.add(m.group(1))
        if cur is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if not m:
            continue
        counts[cur]['instructions'] += 1
        op = m.group(1)
        for mn in MNEMONICS:
            if op.startswith(mn):
                counts[cur][mn] += 1
                total[mn] += 1
    names = demangle(sorted(counts))
    print(f'# {os.path.relpath(lib, ROOT)}: arch {sorted(arch)}, {len(counts)} kernels')
    print('# totals: ' + ', '.join(f'{k} {total[k]}' for k in MNEMONICS if total[k]))
    cols = ('REG', 'SHARED', 'STACK', 'LOCAL', 'instructions') + tuple(k for k in MNEMONICS if total[k])
    print('kernel | ' + ' | '.join(cols))
    for k in sorted(counts, key=lambda k: names[k]):
        u = usage.get(k, {})
        row = [u.get('REG', '?'), u.get('SHARED', '?'), u.get('STACK', '?'), u.get('LOCAL', '?'), str(counts[k]['instructions'])]
        row += [str(counts[k][mn]) for mn in MNEMONICS if total[mn]]
        print(names[k][:110] + ' | ' + ' | '.join(row))
    frames = [k for k in counts if int(usage.get(k, {}).get('STACK', '0')) > 0]
    spilled = sorted(names[k] for k in counts if counts[k]['LDL'] > 0)
    print(f'# {len(frames)} kernels have a stack frame of 8-48 bytes: the vprintf argument buffer of the mbarrier-timeout diagnostic '
          f'(STL only) and, in {len(spilled)} of them, a few values spilled at the register cap and read back (LDL > 0): '
          + ', '.join(n[:40] for n in spilled))


if __name__ == '__main__':
    main()
