#!/usr/bin/env python
"""Turns an .ncu-rep (brought back under gpurun_out/) into the short text summary kept in profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_ncu.txt

Per captured launch: duration, DRAM bytes read/written (the `traffic` of bench.py's roofline object),
L2 hit rate, occupancy, issue utilisation, shared-memory wavefronts / bank conflicts, warp-stall
breakdown and the source lines that execute the most instructions (needs -lineinfo).
"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor']


def main(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for ri in range(2, len(rows)):
        print('=== launch %d: %s' % (ri - 2, rows[ri][hdr.index('Kernel Name')][:100]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('  %-68s %s %s' % (k, rows[ri][i], units[i]))
        st = {h.replace('smsp__pcsamp_warps_issue_stalled_', ''): int(rows[ri][i]) for i, h in enumerate(hdr)
              if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')}
        tot = max(1, sum(st.values()))
        print('  warp stalls: ' + ', '.join('%s %.0f%%' % (k, 100.0 * v / tot)
                                            for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
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
        agg[k] = (agg.get(k, (0, 0, ''))[0] + n, agg.get(k, (0, 0, ''))[1] + sm, r[1])
    tot = max(1, sum(v[0] for v in agg.values()))
    tots = max(1, sum(v[1] for v in agg.values()))
    print('=== source lines by executed warp-instructions (last captured launch)')
    for (f, l), (n, sm, text) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
        print('  %s:%-4d inst %5.2f%%  stall-samples %5.2f%%  %s' % (f, l, 100.0 * n / tot, 100.0 * sm / tots, text.strip()[:90]))


if __name__ == '__main__':
    main(sys.argv[1])
