#!/usr/bin/env python
"""Per-source-line instruction / stall / shared-wavefront breakdown of ONE launch of an .ncu-rep.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep <launch index> [top N]
"""
import csv
import subprocess
import sys


def main(rep, launch, top=30):
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--launch-skip', str(launch),
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    agg, cur, hdr = {}, None, None
    for r in csv.reader(src.splitlines()):
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == 'Function Name' or r[2] != '-':
            continue
        try:
            n = int(r[hdr.index('Instructions Executed')])
            sm = int(r[hdr.index('# Samples')])
            wf = int(r[hdr.index('L1 Wavefronts Shared')])
            wfi = int(r[hdr.index('L1 Wavefronts Shared Ideal')])
        except ValueError:
            continue
        k = (cur, int(r[0]))
        a = agg.get(k, [0, 0, 0, 0, r[1]])
        a[0] += n; a[1] += sm; a[2] += wf; a[3] += wfi
        agg[k] = a
    tot = max(1, sum(v[0] for v in agg.values()))
    tots = max(1, sum(v[1] for v in agg.values()))
    totw = max(1, sum(v[2] for v in agg.values()))
    print('total inst %d  samples %d  shared wavefronts %d (ideal %d)' % (tot, tots, totw, sum(v[3] for v in agg.values())))
    for (f, l), (n, sm, wf, wfi, text) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print('  %s:%-4d inst %5.2f%%  stall %5.2f%%  smem-wf %5.2f%% (x%.2f)  %s' % (
            f, l, 100.0 * n / tot, 100.0 * sm / tots, 100.0 * wf / totw, wf / max(1, wfi), text.strip()[:80]))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 30)
