#!/usr/bin/env python3
"""helpers to read `ncu --page source --csv` dumps: per-opcode and per-region instruction / stall-sample counts"""
import csv, sys
from collections import Counter

def load(path, which=0):
    rows = list(csv.reader(open(path)))
    secs = []
    for r in rows:
        if r and r[0] == 'Kernel Name':
            secs.append([r]); continue
        if secs: secs[-1].append(r)
    sec = secs[which]
    hdr = sec[1]
    data = [r for r in sec[2:] if len(r) == len(hdr)]
    return sec[0][1], hdr, data

if __name__ == '__main__':
    name, hdr, data = load(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    iS, iI, iSrc, iA = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('Address')
    tot = sum(int(r[iI]) for r in data); ts = sum(int(r[iS]) for r in data)
    print(name[:100]); print('total inst', tot, 'samples', ts, 'lines', len(data))
    c = Counter(); s = Counter()
    for r in data:
        t = r[iSrc].strip().split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        c[op] += int(r[iI]); s[op] += int(r[iS])
    for op, n in c.most_common(28):
        print(f'{op:10s} {n:10d} {100*n/tot:5.1f}%  samples {100*s[op]/ts:5.1f}%')
    # dump executed lines with samples
    if len(sys.argv) > 3:
        base = int(data[0][iA], 16)
        for r in data:
            if int(r[iI]) > 0:
                print(f"{int(r[iA],16)-base:05x} {int(r[iI]):9d} {int(r[iS]):6d}  {r[iSrc].strip()[:90]}")
