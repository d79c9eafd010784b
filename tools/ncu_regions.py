#!/usr/bin/env python
"""Stall samples / executed instructions of the LAST launch in an .ncu-rep, grouped by source-line ranges.

    python tools/ncu_regions.py rep file.cuh name:lo-hi [name:lo-hi ...]
"""
import csv
import subprocess
import sys


def main(rep, fname, regions):
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                         capture_output=True, text=True).stdout
    agg, cur = {}, None
    for r in csv.reader(src.splitlines()):
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if len(r) < 8 or r[0] in ('Line No', 'Function Name') or r[2] != '-':
            continue
        try:
            n, sm = int(r[7]), int(r[6])
        except ValueError:
            continue
        k = (cur, int(r[0]))
        a = agg.get(k, (0, 0))
        agg[k] = (a[0] + n, a[1] + sm)
    tot = max(1, sum(v[0] for v in agg.values()))
    tots = max(1, sum(v[1] for v in agg.values()))
    out = {}
    for (f, l), (n, sm) in agg.items():
        name = 'other:' + str(f)
        if f == fname:
            name = 'unassigned'
            for rn, lo, hi in regions:
                if lo <= l <= hi:
                    name = rn
        a = out.get(name, [0, 0])
        a[0] += n
        a[1] += sm
        out[name] = a
    for k, (n, sm) in sorted(out.items(), key=lambda kv: -kv[1][1]):
        print('%-24s stall-samples %5.1f%%   instructions %5.1f%%' % (k, 100.0 * sm / tots, 100.0 * n / tot))


if __name__ == '__main__':
    regs = []
    for a in sys.argv[3:]:
        n, r = a.split(':')
        lo, hi = r.split('-')
        regs.append((n, int(lo), int(hi)))
    main(sys.argv[1], sys.argv[2], regs)
