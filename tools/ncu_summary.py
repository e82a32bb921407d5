#!/usr/bin/env python
"""One line per profiled launch of an ncu report: duration, instructions, active threads per instruction, issue slot use,
DRAM / L2 / local traffic and the top stall reasons.   python tools/ncu_summary.py gpurun_out/x.ncu-rep [--csv]"""
import csv
import subprocess
import sys

KEYS = {
    'gpu__time_duration.sum': 'ms',
    'smsp__inst_executed.sum': 'winst',
    'smsp__thread_inst_executed_per_inst_executed.ratio': 'act',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ%',
    'dram__bytes_read.sum': 'dramR',
    'dram__bytes_write.sum': 'dramW',
    'lts__t_sectors.sum': 'l2sect',
    'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum': 'locLd',
    'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum': 'locSt',
    'launch__registers_per_thread': 'regs',
}
STALLS = 'smsp__average_warps_issue_stalled_'


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index('Kernel Name')
    for r in rows[2:]:
        d = {}
        stalls = []
        for i, h in enumerate(hdr):
            if h in KEYS:
                d[KEYS[h]] = (r[i], units[i])
            if h.startswith(STALLS) and h.endswith('_per_issue_active.ratio'):
                try:
                    stalls.append((float(r[i]), h[len(STALLS):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print(r[name_i][:60])
        print('   ' + '  '.join('%s=%s%s' % (k, v[0], '' if v[1] in ('', 'inst', 'sector', 'register/thread') else v[1]) for k, v in d.items()))
        print('   stalls: ' + ', '.join('%s %.2f' % (n, v) for v, n in stalls[:6]))


if __name__ == '__main__':
    main()
