#!/usr/bin/env python
"""profiles/r02_traffic.json from `ncu --set full` captures of tools/prof_stage.py (256 MiB): DRAM read + write bytes of the
LAST launch of every kernel in a report, per input byte.  bench.py scales these for `roofline.traffic`.

  python tools/make_traffic.py gpurun_out/r02_bpe_full.ncu-rep gpurun_out/r02_uni_full.ncu-rep gpurun_out/r02_pipe_full2.ncu-rep
(the BPE report holds two passes: cold word cache, then warm -- the last launch is the warm one)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INPUT = 256 << 20
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main():
    out = {}
    for rep in sys.argv[1:]:
        txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        ni, ri, wi = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        uni = 'uni' in os.path.basename(rep)
        for r in rows[2:]:
            name = r[ni].split('(')[0].replace('void ', '')
            name = {'ak_resolve_kernel<0>': 'ak_resolve_kernel<bpe>', 'ak_resolve_kernel<1>': 'ak_resolve_kernel<unigram>',
                    'ak_words_kernel<0>': 'ak_words_kernel', 'ak_words_kernel<1>': 'ak_words_kernel<unigram>',
                    'ak_emit_kernel<int>': 'ak_emit_kernel<unigram run>' if uni else 'ak_emit_kernel'}.get(name, name)
            b = float(r[ri].replace(',', '')) * UNIT[units[ri]] + float(r[wi].replace(',', '')) * UNIT[units[wi]]
            out[name] = {'dram_bytes_per_input_byte': round(b / INPUT, 4), 'dram_bytes': int(b), 'input_bytes': INPUT,
                         'report': os.path.basename(rep) + ' (last launch of the kernel in the capture)'}
    with open(os.path.join(ROOT, 'profiles', 'r02_traffic.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in sorted(out.items()):
        print('%-32s %.3f' % (k, v['dram_bytes_per_input_byte']))


if __name__ == '__main__':
    main()
