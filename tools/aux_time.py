#!/usr/bin/env python
"""Device time of the kernels either side of the path at 1 GiB (CUDA events around the public Engine calls, 3 runs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tools')]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import akshar_b200 as A  # noqa: E402
from akshar_b200 import _lib as C  # noqa: E402
import synth_corpus as sc  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    eng = A.Engine(0)
    data, off = sc.Corpus('hindi', 20261018).generate(mb << 20)
    gib = data.size / 2 ** 30
    b = eng.put((torch.from_numpy(data), torch.from_numpy(off)))
    norm = eng.normalize_batch(b)
    print('rows %d, bytes %d' % (b.n_rows, b.n_bytes))
    # file -> rows: the same text as a file (rows joined by newlines, some padding spaces)
    f = eng.join_rows(b)
    t = timed(lambda: eng.lines_batch(f, f.numel(), row_capacity=b.n_rows + 16))
    print('lines (file bytes -> rows)          %.2f ms per GiB (includes one read-back of the totals)' % (t / (f.numel() / 2 ** 30)))
    t = timed(lambda: eng.join_rows(b))
    print('join rows                           %.2f ms per GiB' % (t / gib))
    for rule, name in ((C.WORDS_HINDI, 'hindi rule'), (C.WORDS_SPLIT, 'split rule')):
        t = timed(lambda: eng.word_tokenize_batch(norm, rule=rule))
        print('word tokenizer (%s)         %.2f ms per GiB (includes one read-back)' % (name, t / (norm.n_bytes / 2 ** 30)))
    cl, ru = eng.segment_batch(norm, clusters=True, runs=True)
    stats = torch.empty((b.n_rows, 5), dtype=torch.int32, device=eng.device)

    def comp():
        eng.lib.akshar_composition_batch(eng._h, norm.data.data_ptr(), norm.offsets.data_ptr(), norm.n_rows, cl.splits.data_ptr(),
                                         ru.values.data_ptr(), ru.extra.data_ptr(), ru.splits.data_ptr(), stats.data_ptr(), eng._stream())
    t = timed(comp)
    print('composition counts                  %.2f ms per GiB' % (t / (norm.n_bytes / 2 ** 30)))
    t = timed(lambda: eng.merge_clusters_batch(norm, cl, C.MERGE_AKSHARA))
    print('cluster merge (akshara)             %.2f ms per GiB (includes one read-back)' % (t / (norm.n_bytes / 2 ** 30)))
    tk = A.aksharTokenizer(os.path.join(ROOT, 'tests', 'golden', 'models', 'spm24k.model'), 'sentencepiece')
    ids = tk._eng.encode_unigram_batch(tk._eng.put(norm))
    t = timed(lambda: tk._eng.decode_batch(ids, None, 1))
    print('decode (Unigram, %d ids -> text)  %.2f ms per GiB of text (includes one read-back)' % (ids.values.numel(), t / (norm.n_bytes / 2 ** 30)))


if __name__ == '__main__':
    main()
