#!/usr/bin/env python
"""Golden vectors for the tokenizer with clean_hinglish=False (text outside normalize_text's closed alphabet reaches the
models), recorded from the UNMODIFIED reference in the dev container:  python tools/make_golden_raw.py
-> tests/golden/reference_vectors_raw.json.gz"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
sys.path.insert(0, '/root/reference/src')
import synth_corpus as sc  # noqa: E402
from make_golden import HANDPICKED, MODELS, check_pins  # noqa: E402


def hf_fuzz(n):
    """rows for what only HF's side of the BPE path sees (scripts/train_bpe.py:71,80): compatibility characters, marks to
    reorder / compose, code points newer than HF's Unicode tables, and the added-token syntax in the raw text"""
    import random
    import unicodedata
    rng = random.Random(20261018)
    compat = [c for c in range(0x110000) if not 0xD800 <= c <= 0xDFFF and
              unicodedata.normalize('NFKD', chr(c)) != unicodedata.normalize('NFD', chr(c))]
    marks = list(range(0x300, 0x370)) + [0x93c, 0x94d, 0x5bc, 0x5c1, 0x3099, 0x309a, 0x9be, 0x9d7, 0x1161, 0x11a8, 0x9fe, 0x11935,
                                         0x11930, 0x11938, 0x1e08f, 0x10f46]
    words = ['hello', 'yaar', 'kya', '\u0928\u092e\u0938\u094d\u0924\u0947', '\u0915', '\u0937', 'x', '2', ' ', ' ', ' ', '.', ',', '<', '>', '/',
             '<s>', '</s>', '<pad>', '<unk>', '<mask>', '<s', 's>', '<S>', '< s>', '<<s>>', '\u00e9', 'e\u0301', '\U0001F600', '\ufb01',
             '\u2460', 'x\u00b2', '\uff26\uff55\uff4c\uff4c', '\u2026', '\u2122', '\u00bd', '\u3000', '\u00a0', '\ufdfa', '\u1e9b\u0323', '\u212b',
             '\uac01', '\u1100\u1161\u11a8', '\n', '\t']
    out = []
    for _ in range(n):
        parts = []
        for _ in range(rng.randint(1, 9)):
            r = rng.random()
            if r < 0.55:
                parts.append(rng.choice(words))
            elif r < 0.8:
                parts.append(chr(rng.choice(compat)))
            else:
                parts.append(chr(rng.choice(marks)))
        out.append(''.join(parts))
    return out


def main():
    pins = check_pins()
    from akshar.tokenizer import aksharTokenizer
    from akshar import normalize as RN
    inputs = [s for s in HANDPICKED] + sc.adversarial(1500, 17, 40) + sc.Corpus('social', 19).lines(60000)
    inputs += hf_fuzz(1200)
    inputs += ['naïve café 😀👍🏽 ok', 'Ångström ﬁ ① x²', '日本語 テキスト', 'mixed nbsp　wide', 'ạ́b', 'tab\there\nnewline']
    tb = aksharTokenizer(os.path.join(MODELS, 'bpe24k.json'), 'bpe', clean_hinglish=False)
    tu = aksharTokenizer(os.path.join(MODELS, 'spm24k.model'), 'sentencepiece', clean_hinglish=False)
    tb2 = aksharTokenizer(os.path.join(MODELS, 'bpe24k.json'), 'bpe', normalize_roman=False, clean_hinglish=False)
    from tokenizers import normalizers
    nfkc = normalizers.NFKC()
    rows = []
    for s in inputs:
        rows.append({'in': s, 'norm_nc': RN.normalize_text(s, clean_hinglish=False), 'ids_bpe24k': tb.encode(s),
                     'ids_spm24k': tu.encode(s), 'ids_bpe24k_raw': tb2.encode(s), 'hf_nfkc': nfkc.normalize_str(s)})
    path = os.path.join(ROOT, 'tests', 'golden', 'reference_vectors_raw.json.gz')
    with gzip.GzipFile(path, 'wb', mtime=0) as f:
        f.write(json.dumps({'pins': pins, 'rows': rows}, ensure_ascii=False, separators=(',', ':')).encode('utf-8'))
    print('wrote', path, os.path.getsize(path), len(rows), 'rows')


if __name__ == '__main__':
    main()
