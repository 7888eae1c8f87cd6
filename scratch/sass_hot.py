#!/usr/bin/env python
"""Hot-spot summary of an `ncu --page source --csv` export: opcode mix (executed / stall samples) and the most sampled SASS lines.
usage: sass_hot.py file.csv [ntop]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h) and r[0].startswith("0x")]
iS, iE, isrc = h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed'), h.index('Source')
I = lambda s: int(s) if s.strip() else 0
tot = sum(I(r[iS]) for r in data); totE = sum(I(r[iE]) for r in data)
print('total samples', tot, 'warp instr executed', totE, 'SASS lines', len(data))
ex = defaultdict(int); sm = defaultdict(int)
for r in data:
    t = r[isrc].split()
    if not t: continue
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    op = op.split('.')[0].rstrip(';')
    ex[op] += I(r[iE]); sm[op] += I(r[iS])
print('opcode mix:')
for k, v in sorted(ex.items(), key=lambda x: -x[1])[:24]:
    print('  %-10s %6.2f%% exec  %6.2f%% samples' % (k, 100 * v / max(totE, 1), 100 * sm[k] / max(tot, 1)))
print('top sampled lines:')
for r in sorted(data, key=lambda r: -I(r[iS]))[:ntop]:
    print('  %5d %7d  %s' % (I(r[iS]), I(r[iE]), r[isrc].strip()[:100]))
