#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples for launch K of an .ncu-rep (needs --import-source on).

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep [launch_index] [top_n]
"""
import csv
import subprocess
import sys


def main(rep, k=0, top=40):
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, idx, cur = None, -1, []
    launches = []
    for r in rows:
        if r and r[0] == 'Address':
            hdr = r
            cur = []
            launches.append(cur)
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        cur.append(dict(zip(hdr, r)))
    L = launches[k]
    tot = sum(int(d['# Samples'] or 0) for d in L) or 1
    ins = sum(int(d['Instructions Executed'] or 0) for d in L) or 1
    print('launch %d: %d SASS instructions, %d stall samples, %d warp-instructions executed' % (k, len(L), tot, ins))
    keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {h: sum(int(d[h] or 0) for d in L) for h in keys}
    print('stalls: ' + ', '.join('%s %.1f%%' % (h[6:], 100.0 * v / tot) for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(L)), key=lambda i: -int(L[i]['# Samples'] or 0))[:top]
    for i in sorted(order):
        d = L[i]
        s = int(d['# Samples'] or 0)
        why = sorted(((int(d[h] or 0), h[6:]) for h in keys), reverse=True)[:2]
        print('%5d %5.2f%% inst %5.2f%%  %-70s %s' % (i, 100.0 * s / tot, 100.0 * int(d['Instructions Executed'] or 0) / ins,
                                                    d['Source'][:70], ' '.join('%s:%d' % (n, v) for v, n in why if v)))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
