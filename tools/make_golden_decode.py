#!/usr/bin/env python
"""Generate tests/golden/decode_fuzz.json.gz by running the UNMODIFIED reference tokenizer's decode / detokenize
(/root/reference/src/akshar/tokenizer.py:195-246, i.e. sentencepiece 0.2.1 DecodeIds and tokenizers 0.22.2 decode) over
seeded id rows the encoders would never produce: byte pieces in and out of order, control / unknown / special ids, ids the
vocabulary does not use, pieces that are white space only.  They pin oracle.unigram_decode / bpe_decode / *_detokenize and,
through them, the on-device decode (ak_decode.cuh).

  python tools/make_golden_decode.py
"""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, os.path.join(ROOT, 'tools'))
sys.path.insert(0, os.path.join(REF, 'src'))
MODELS = os.path.join(ROOT, 'tests', 'golden', 'models')

from make_golden import check_pins  # noqa: E402


def id_rows(rng, size, n, interesting, hi_extra=0):
    rows = [[], [0], [1], [2], [3], [1, 2], [size - 1]]
    for _ in range(n):
        k = int(rng.integers(0, 24))
        mode = rng.random()
        if mode < 0.35:
            r = rng.integers(0, size + hi_extra, size=k).tolist()
        elif mode < 0.75:
            r = [int(interesting[int(j)]) for j in rng.integers(0, len(interesting), size=k)]
        else:
            r = []
            for _ in range(k):
                r.append(int(interesting[int(rng.integers(len(interesting)))]) if rng.random() < 0.5 else int(rng.integers(0, size)))
        rows.append([int(x) for x in r])
    return rows


def main():
    pins = check_pins()
    from akshar.tokenizer import aksharTokenizer
    rng = np.random.default_rng(20261018)
    out = {'pins': pins}
    # ---- SentencePiece
    for name in ('spm24k', 'spm_corpus'):
        tk = aksharTokenizer(os.path.join(MODELS, name + '.model'), 'sentencepiece')
        sp = tk.model
        size = sp.GetPieceSize()
        us = sp.PieceToId('▁')
        byte_ids = [sp.PieceToId('<0x%02X>' % b) for b in range(256)]
        # well-formed and malformed UTF-8 as byte pieces, next to the pieces that matter at the start of a text
        seqs = ['क', 'é', '\U0001F600', 'A', ' ', '▁', ' ', '　']
        interesting = [0, 1, 2, 3, us, us, 260, 261, 300, 301]
        for s in seqs:
            interesting += [byte_ids[b] for b in s.encode('utf-8')]
        interesting += [byte_ids[b] for b in (0x80, 0xBF, 0xC0, 0xC1, 0xC2, 0xE0, 0xA0, 0x9F, 0xED, 0xF0, 0x90, 0x8F, 0xF4, 0xF5, 0xFF, 0x00, 0x7F, 0x20)]
        rows = id_rows(rng, size, 1500, interesting)
        # hand-made: leading-space rule, runs of bytes cut by pieces and by the end of the row
        b = lambda x: byte_ids[x]
        rows += [[b(0xE0), b(0xA4), b(0x95)], [b(0xE0), b(0xA4)], [b(0xA4), b(0x95)], [b(0xE0), 300, b(0xA4), b(0x95)],
                 [b(0xF0), b(0x9F), b(0x98), b(0x80)], [b(0xC0), b(0x80)], [b(0xED), b(0xA0), b(0x80)], [b(0xF4), b(0x90), b(0x80), b(0x80)],
                 [b(0x41), b(0xFF), b(0x42)], [0, 300, 0], [1, 300, 2, 3], [b(0xE2), b(0x96), b(0x81), 300], [us, 300], [us, us, 300],
                 [1, us, 300], [us], [us, us], [300, us], [300, us, us, 301], [1, 1, us, 2, 300], [b(0x20), 300],
                 [b(0xE0), b(0xA4), b(0x95), b(0xE0)], [b(0xC2), b(0xA0)], [b(0xE0), b(0x80), b(0x80)], [b(0xF0), b(0x8F), b(0xBF), b(0xBF)],
                 [b(0xF0), b(0x9F), b(0x98)], [b(0x9F), b(0x98), b(0x80)], [b(0xE0), b(0xA4), b(0x95)] * 5]
        dec = [tk.decode(r) for r in rows]
        pieces = [[sp.IdToPiece(i) for i in r] for r in rows]
        det = [tk.detokenize(p) for p in pieces]
        out[name] = {'ids': rows, 'decode': dec, 'detokenize': det}
    # ---- HF BPE
    for name in ('bpe24k', 'bpe_corpus'):
        tk = aksharTokenizer(os.path.join(MODELS, name + '.json'), 'bpe')
        hf = tk.model
        size = hf.get_vocab_size()
        vocab = hf.get_vocab()
        interesting = [0, 1, 2, 3, 4, 5, 6, 7]
        interesting += [i for t, i in vocab.items() if t.startswith('#') or 'Ġ' in t][:40]
        interesting += [int(x) for x in rng.integers(0, size, size=30)]
        rows = id_rows(rng, size, 1500, interesting, hi_extra=50)
        dec = [tk.decode(r) for r in rows]
        # detokenize works on token strings: ids without a token cannot be part of `tokenize` output
        rows_ok = [[i for i in r if hf.id_to_token(i) is not None] for r in rows]
        det = [tk.detokenize([hf.id_to_token(i) for i in r]) for r in rows_ok]
        out[name] = {'ids': rows, 'decode': dec, 'ids_detok': rows_ok, 'detokenize': det}
    # detokenize on token strings no vocabulary holds (tokenizer.py:236-246 is plain string work)
    toks = [['##a', '##b'], ['a', '##b', 'c'], ['Ġa', 'Ġ'], ['Ġ', '##x'], ['a', '##', 'b'], ['a', '###'], [' ', 'a', ' '],
            ['▁', '▁a', '▁'], ['　a '], [], ['##'], ['a', 'Ġ##b']]
    tb = aksharTokenizer(os.path.join(MODELS, 'bpe24k.json'), 'bpe')
    tu = aksharTokenizer(os.path.join(MODELS, 'spm24k.model'), 'sentencepiece')
    out['detokenize_strings'] = {'tokens': toks, 'bpe': [tb.detokenize(t) for t in toks], 'sentencepiece': [tu.detokenize(t) for t in toks],
                                 'akshar': [aksharTokenizer().detokenize(t) for t in toks]}
    path = os.path.join(ROOT, 'tests', 'golden', 'decode_fuzz.json.gz')
    with gzip.GzipFile(path, 'wb', mtime=0) as f:
        f.write(json.dumps(out, ensure_ascii=False, sort_keys=True).encode('utf-8'))
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
