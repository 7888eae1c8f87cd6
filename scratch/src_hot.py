#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump: executed instructions by opcode and the hottest stall PCs."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; body = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit()]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iE]) for r in body); tots = sum(int(r[iSm]) for r in body)
ops = collections.Counter(); samp = collections.Counter()
for r in body:
    src = r[iS].strip()
    op = src.split()[0]
    if op.startswith('@'): op = src.split()[1]
    op = op.split('.')[0]
    ops[op] += int(r[iE]); samp[op] += int(r[iSm])
print("total inst executed %d, samples %d" % (tot, tots))
for op, n in ops.most_common(22):
    print("  %-10s %6.2f%% inst   %6.2f%% samples" % (op, 100.0*n/tot, 100.0*samp[op]/max(tots,1)))
# regions: cumulative executed instruction by address blocks of 64 instrs to see loops
if len(sys.argv) > 2:
    print("hottest instructions by stall samples:")
    for r in sorted(body, key=lambda r: -int(r[iSm]))[:int(sys.argv[2])]:
        print("  %6s samples %9s exec  %s" % (r[iSm], r[iE], r[iS].strip()[:90]))
