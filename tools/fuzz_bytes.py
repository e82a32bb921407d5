#!/usr/bin/env python
"""Arbitrary bytes through every entry point, one call at a time with AKSHAR_DEBUG_SYNC=1 (names the kernel that faults)."""
import os
import sys
os.environ['AKSHAR_DEBUG_SYNC'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tools')]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import akshar_b200 as A  # noqa: E402
from akshar_b200 import _lib as C  # noqa: E402


def main():
    rng = np.random.default_rng(99)
    models = os.path.join(ROOT, 'tests', 'golden', 'models')
    eng = A.Engine(0)
    tb = A.aksharTokenizer(os.path.join(models, 'bpe24k.json'), 'bpe')
    tu = A.aksharTokenizer(os.path.join(models, 'spm24k.model'), 'sentencepiece')
    pools = [np.arange(256, dtype=np.uint8),
             np.array([0x80, 0xBF, 0xC0, 0xC2, 0xE0, 0xA4, 0xA5, 0xBC, 0x8D, 0xED, 0xF0, 0xF4, 0xFF, 0x20, 0x61, 0x0A, 0x3C, 0x73, 0x3E], dtype=np.uint8)]
    for pi, pool in enumerate(pools):
        for n_rows, max_len in ((3000, 200), (20, 20000), (1, 300000)):
            lens = rng.integers(0, max_len + 1, size=n_rows)
            off = np.zeros(n_rows + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            data = pool[rng.integers(0, pool.size, size=int(off[-1]))]
            host = (torch.from_numpy(data), torch.from_numpy(off))

            def step(name, fn):
                print('pool', pi, 'rows', n_rows, name, flush=True)
                try:
                    fn()
                except Exception as ex:
                    print('   raised:', str(ex)[:200], flush=True)
                torch.cuda.synchronize()
            b = eng.put(host)
            step('normalize', lambda: eng.normalize_batch(b, check=False))
            step('normalize raw', lambda: eng.normalize_batch(b, clean_hinglish=False, check=False))
            step('segment', lambda: eng.segment_batch(b, clusters=True, runs=True, check=False))
            step('masks', lambda: eng.segment_masks(b, clusters=True, runs=True, check=False))
            step('norm+seg', lambda: eng.normalize_segment_batch(b, check=False))
            step('words hindi', lambda: eng.word_tokenize_batch(b, rule=C.WORDS_HINDI, row_flags=True))
            step('words split', lambda: eng.word_tokenize_batch(b, rule=C.WORDS_SPLIT, row_flags=True))
            step('lines', lambda: eng.lines_batch(b.data, b.n_bytes))
            for tk, kind, nm in ((tb, 0, 'bpe'), (tu, 1, 'unigram')):
                for mode in (C.MODE_TILES, C.MODE_ROWS):
                    step('%s mode %d' % (nm, mode), lambda: tk._eng.tokenizer_encode_batch(host, kind, mode=mode, check=False))
                    step('%s raw mode %d' % (nm, mode), lambda: tk._eng.tokenizer_encode_batch(host, kind, clean_hinglish=False, mode=mode, check=False))
    print('survived')


if __name__ == '__main__':
    main()
