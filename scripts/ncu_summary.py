#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel, --set full) into a small text file for profiles/."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'dram__bytes_read.sum.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'launch__grid_size',
        'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active']


def main(rep, out):
    raw = subprocess.check_output(['ncu', '-i', rep, '--page', 'raw', '--csv'], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"== {d.get('Kernel Name', '?')[:110]}")
        for k in KEYS:
            if k in d:
                lines.append(f'{k:78s} {d[k]:>16s} {units[hdr.index(k)]}')
        lines.append('-- warp stall reasons (warps per issue-active cycle)')
        for h in hdr:
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                name = re.sub(r'smsp__average_warps_issue_stalled_|_per_issue_active.ratio', '', h)
                if float(d[h] or 0) >= 0.01:
                    lines.append(f'   {name:28s} {float(d[h]):8.3f}')
    src = subprocess.check_output(['ncu', '-i', rep, '--page', 'source', '--csv'], text=True)
    srows = list(csv.reader(io.StringIO(src)))
    h2 = srows[1]
    ia, ie = h2.index('Source'), h2.index('Instructions Executed')
    op = defaultdict(int)
    tot = 0
    for r in srows[2:]:
        if len(r) <= ie:
            continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia])
        o = m.group(2).split('.')[0] if m else '?'
        op[o] += int(r[ie])
        tot += int(r[ie])
    lines.append(f'-- SASS mix (warp instructions executed, total {tot})')
    for o, c in sorted(op.items(), key=lambda x: -x[1])[:18]:
        lines.append(f'   {o:10s} {c:12d} {100 * c / max(tot, 1):5.1f}%')
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines[:60]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
