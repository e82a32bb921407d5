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


def main():
    pins = check_pins()
    from akshar.tokenizer import aksharTokenizer
    from akshar import normalize as RN
    inputs = [s for s in HANDPICKED] + sc.adversarial(1500, 17, 40) + sc.Corpus('social', 19).lines(60000)
    inputs += ['naïve café 😀👍🏽 ok', 'Ångström ﬁ ① x²', '日本語 テキスト', 'mixed nbsp　wide', 'ạ́b', 'tab\there\nnewline']
    tb = aksharTokenizer(os.path.join(MODELS, 'bpe24k.json'), 'bpe', clean_hinglish=False)
    tu = aksharTokenizer(os.path.join(MODELS, 'spm24k.model'), 'sentencepiece', clean_hinglish=False)
    tb2 = aksharTokenizer(os.path.join(MODELS, 'bpe24k.json'), 'bpe', normalize_roman=False, clean_hinglish=False)
    rows = []
    for s in inputs:
        rows.append({'in': s, 'norm_nc': RN.normalize_text(s, clean_hinglish=False), 'ids_bpe24k': tb.encode(s),
                     'ids_spm24k': tu.encode(s), 'ids_bpe24k_raw': tb2.encode(s)})
    path = os.path.join(ROOT, 'tests', 'golden', 'reference_vectors_raw.json.gz')
    with gzip.GzipFile(path, 'wb', mtime=0) as f:
        f.write(json.dumps({'pins': pins, 'rows': rows}, ensure_ascii=False, separators=(',', ':')).encode('utf-8'))
    print('wrote', path, os.path.getsize(path), len(rows), 'rows')


if __name__ == '__main__':
    main()
