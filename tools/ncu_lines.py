#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an .ncu-rep captured with --import-source on
(kernels compiled with -lineinfo).  usage: python tools/ncu_lines.py REPORT.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur = None
    func = None
    agg = collections.defaultdict(lambda: [0, 0, 0, ''])
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            func = r[1]
            continue
        if cur is None or not r[0].strip().isdigit() or len(r) < 9:
            continue
        try:
            inst, smp, tinst = int(r[7]), int(r[6]), int(r[8])
        except ValueError:
            continue
        a = agg[(func, cur, int(r[0]))]
        a[0] += inst
        a[1] += smp
        a[2] += tinst
        a[3] = r[1]
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[1] for a in agg.values()) or 1
    byfn = collections.defaultdict(lambda: [0, 0])
    for (fn, f, l), a in agg.items():
        byfn[(fn, f)][0] += a[0]
        byfn[(fn, f)][1] += a[1]
    print('total warp instructions %d, samples %d' % (tot, tots))
    for (fn, f), a in sorted(byfn.items(), key=lambda kv: -kv[1][0]):
        print('  %-28s %-60s inst %5.1f%%  samples %5.1f%%' % (f, (fn or '')[:60], 100 * a[0] / tot, 100 * a[1] / tots))
    for (fn, f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
        print('%5.2f%% inst %5.2f%% smp thr/inst %4.1f  %s:%d  %s' % (100 * a[0] / tot, 100 * a[1] / tots, a[2] / max(a[0], 1), f, l,
                                                                       a[3].strip()[:100]))


if __name__ == '__main__':
    main()
