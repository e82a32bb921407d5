#!/usr/bin/env python
"""One pass of each stage over MB MiB of synthetic text, for ncu:  python tools/prof_stage.py [MB] [pipeline|bpe|unigram]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tools')]
import torch  # noqa: E402

import akshar_b200 as A  # noqa: E402
import synth_corpus as sc  # noqa: E402


def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    what = sys.argv[2] if len(sys.argv) > 2 else 'pipeline'
    kind = {'pipeline': 'social', 'bpe': 'hinglish', 'unigram': 'hindi'}[what]
    data, off = sc.Corpus(kind, 20261018).generate(mb << 20)
    models = os.path.join(ROOT, 'tests', 'golden', 'models')
    if what == 'pipeline':
        eng = A.Engine(0)
        b = eng.put((torch.from_numpy(data), torch.from_numpy(off)))
        for _ in range(2):
            eng.normalize_segment_batch(b)
    else:
        tk = A.aksharTokenizer(os.path.join(models, 'bpe24k.json' if what == 'bpe' else 'spm24k.model'), 'bpe' if what == 'bpe' else 'sentencepiece')
        b = tk._eng.put((torch.from_numpy(data), torch.from_numpy(off)))
        for _ in range(2):
            tk._eng.tokenizer_encode_batch(b, 0 if what == 'bpe' else 1)
    torch.cuda.synchronize()
    print('done')


if __name__ == '__main__':
    main()
