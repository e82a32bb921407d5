#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference in the dev container.

  python tools/make_golden.py [--train]

* imports the reference package from /root/reference/src (read-only) and records, for a fixed seeded set of
  inputs, the outputs of every hot-path function;
* with --train, first re-creates the four model files under tests/golden/models/ by running the reference's
  own scripts/train_bpe.py and scripts/train_spm.py (BPE-24k on 30 MB synthetic Hinglish, Unigram-24k on 30 MB
  synthetic Hindi, and both on data/corpus.txt; train_spm.py needs --vocab-size 400 there).

The reference cannot travel to the GPU box, the vectors can: tests compare oracle/ and the CUDA path to them.
Pinned engine versions are stored in the fixture and asserted here.
"""
import gzip
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, os.path.join(ROOT, 'tools'))
sys.path.insert(0, os.path.join(REF, 'src'))

import synth_corpus as sc  # noqa: E402

PINS = {'regex': '2026.3.32', 'tokenizers': '0.22.2', 'sentencepiece': '0.2.1', 'unicodedata': '15.0.0'}
MODELS = os.path.join(ROOT, 'tests', 'golden', 'models')

# expectations written in the reference's own tests / executed notebook cells
HANDPICKED = [
    "Hello World", "नमस्ते", "hello नमस्ते world", "yaaaaar", "bohoooot", "Heyyy यार kya HAAL hai", "École",
    "İstanbul", "Kelvin", "a😀a😀a", "<s>hi</s>", "100000 रु", "ＡＢＣ", "नमस्तेेेे", "a_b@c#d",
    "aaj मौसम बहुत अच्छा है", "क्षेत्रे धर्मक्षेत्रे", "अच्छा", "क्‍या", "क्‌ष", "\r\n", "👨‍👩‍👧", "🇮🇳🇮",
    "मौसम", "च्छा", "हूँ", "", " ", "   ", "\n\n\n", "\r\r\r", "!!!", "ााा", "<s> hi </s>", "zzq xq", "5", ".",
    "aaj मौसम", "...", "123", "a\tb", "a b", "मैं स्कूल जा रहा हूँ", "yaar kya haal hai",
]
SIG_WORDS = ["nahi", "nahii", "nahee", "Nahiii", "achha", "bhookh", "chhota", "kkhh", "khoo", "acha", "ee\n", "Σ",
             "ΑΣ", "ΑΣa", "aΣ.", "İ", "yaaaar", "KHAANA", "dhoodh", "phool", ""]


def check_pins():
    import unicodedata
    import regex
    import tokenizers
    import sentencepiece
    got = {'regex': regex.__version__, 'tokenizers': tokenizers.__version__, 'sentencepiece': sentencepiece.__version__,
           'unicodedata': unicodedata.unidata_version}
    assert got == PINS, got
    return got


def train():
    tmp = '/tmp/akshar_train'
    os.makedirs(tmp, exist_ok=True)
    os.makedirs(MODELS, exist_ok=True)
    for kind, seed, name in (('hinglish', 101, 'hinglish_train.txt'), ('hindi', 102, 'hindi_train.txt')):
        d, o = sc.Corpus(kind, seed).generate(30 << 20)
        b = d.tobytes()
        with open(os.path.join(tmp, name), 'wb') as f:
            for i in range(len(o) - 1):
                f.write(b[o[i]:o[i + 1]] + b'\n')
    shutil.copy(os.path.join(REF, 'data', 'corpus.txt'), os.path.join(tmp, 'corpus.txt'))
    run = lambda *a: subprocess.check_call([sys.executable] + list(a), cwd=tmp)
    run(os.path.join(REF, 'scripts', 'train_bpe.py'), 'hinglish_train.txt', '--output', 'bpe24k.json')
    run(os.path.join(REF, 'scripts', 'train_spm.py'), 'hindi_train.txt', '--output', 'spm24k')
    run(os.path.join(REF, 'scripts', 'train_bpe.py'), 'corpus.txt', '--output', 'bpe_corpus.json')
    run(os.path.join(REF, 'scripts', 'train_spm.py'), 'corpus.txt', '--output', 'spm_corpus', '--vocab-size', '400')
    for f in ('bpe24k.json', 'spm24k.model', 'bpe_corpus.json', 'spm_corpus.model'):
        shutil.copy(os.path.join(tmp, f), os.path.join(MODELS, f))


def main():
    pins = check_pins()
    if '--train' in sys.argv:
        train()
    from akshar import normalize as RN, segment as RS, features as RF
    from akshar.tokenizer import aksharTokenizer

    inputs = list(HANDPICKED)
    with open(os.path.join(REF, 'data', 'corpus.txt'), encoding='utf-8') as f:
        inputs += [l.strip() for l in f if l.strip()]
    inputs += sc.adversarial(1200, 7, 40)
    # seeded corpus lines: >= 1000 rows per model of text that is in-vocabulary for the 24k models (the lexicon is fixed,
    # tools/synth_corpus.py LEXICON_SEED; the corpus seed only draws sentences), plus social noise
    for kind, nbytes in (('hinglish', 200000), ('hindi', 290000), ('social', 60000)):
        inputs += sc.Corpus(kind, 9).lines(nbytes)
    # a few long rows for the chunked lanes
    inputs.append(' '.join(sc.Corpus('social', 10).lines(30000)))
    inputs.append(''.join(sc.adversarial(300, 8, 40)))

    toks = {
        'bpe24k': aksharTokenizer(os.path.join(MODELS, 'bpe24k.json'), 'bpe'),
        'spm24k': aksharTokenizer(os.path.join(MODELS, 'spm24k.model'), 'sentencepiece'),
        'bpe_corpus': aksharTokenizer(os.path.join(MODELS, 'bpe_corpus.json'), 'bpe'),
        'spm_corpus': aksharTokenizer(os.path.join(MODELS, 'spm_corpus.model'), 'sentencepiece'),
    }
    fallback = aksharTokenizer()
    rows = []
    for s in inputs:
        norm = RN.normalize_text(s)
        r = {
            'in': s,
            'nfc': RN.normalize_unicode(s),
            'sem': RN.semantic_normalize(s),
            'filt': RN.filter_garbage(s),
            'elong': RN.remove_elongations(s),
            'norm': norm,
            'norm_nr': RN.normalize_text(s, normalize_roman=False),
            'norm_nc': RN.normalize_text(s, clean_hinglish=False),
            'norm_raw': RN.normalize_text(s, normalize_roman=False, clean_hinglish=False),
            'seg_raw': RS.segment_akshars(s),
            'seg': RS.segment_akshars(norm),
            'seg_m': RS.segment_akshars(norm, matras=True),
            'segm_raw': RS.segment_akshars(s, matras=True),
            'cs_raw': RS.detect_code_switches(s),
            'cs': RS.detect_code_switches(norm),
            'comp': RS.analyze_text_composition(norm),
            'tokenize': fallback.tokenize(s),
        }
        for name, tk in toks.items():
            r['ids_' + name] = tk.encode(s)
        r['pieces_spm24k'] = toks['spm24k'].tokenize(s)
        r['pieces_bpe24k'] = toks['bpe24k'].tokenize(s)
        r['dec_bpe24k'] = toks['bpe24k'].decode(r['ids_bpe24k'])
        r['dec_spm24k'] = toks['spm24k'].decode(r['ids_spm24k'])
        # detokenize (tokenizer.py:221-246), tokenize(return_metadata) / explain values (tokenizer.py:123-165, 248-276)
        r['detok_bpe24k'] = toks['bpe24k'].detokenize(r['pieces_bpe24k'])
        r['detok_spm24k'] = toks['spm24k'].detokenize(r['pieces_spm24k'])
        r['detok_akshar'] = fallback.detokenize(r['tokenize'])
        r['meta'] = fallback.tokenize(s, return_metadata=True)
        r['explain_bpe24k'] = toks['bpe24k'].explain(s)
        # word tokenizers (segment.py:239-401)
        r['words_hi'] = RS.word_tokenize_hindi(s)
        r['words_sa'] = RS.word_tokenize_sanskrit(s)
        r['words_auto'] = RS.word_tokenize(s)
        # the feature wrappers that are functions of the cluster boundaries (features.py:28-55, 173-206)
        r['feat_akshara'] = RF.akshara_level_tokenization(s)
        r['feat_nukta'] = RF.preserve_nukta(s)
        rows.append(r)
    sig = {w: RN.roman_phonetic_signature(w) for w in SIG_WORDS}
    words = set()
    for s in inputs[:2000]:
        words.update(s.split()[:3])
    for w in sorted(words):
        sig[w] = RN.roman_phonetic_signature(w)
    ids = {chr(c): RS.identify_script(chr(c)) for c in list(range(0, 0x3000)) + [0x3000, 0xFF21, 0x1D7CE, 0x1F600]
           if not 0xD800 <= c <= 0xDFFF}
    out = {'pins': pins, 'rows': rows, 'signature': sig, 'identify_script': ids,
           'vocab_size': {k: v.vocab_size() for k, v in toks.items()}}
    # BASELINE.json configs[0] runs over the reference's 1.5 KB sample corpus: kept as a data fixture (bench.py's config-1 line)
    shutil.copyfile(os.path.join(REF, 'data', 'corpus.txt'), os.path.join(ROOT, 'tests', 'golden', 'corpus.txt'))
    path = os.path.join(ROOT, 'tests', 'golden', 'reference_vectors.json.gz')
    with gzip.GzipFile(path, 'wb', mtime=0) as f:
        f.write(json.dumps(out, ensure_ascii=False, separators=(',', ':')).encode('utf-8'))
    print('wrote', path, os.path.getsize(path), 'bytes;', len(rows), 'rows')


if __name__ == '__main__':
    main()
