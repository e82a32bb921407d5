"""The oracle (oracle/akshar_oracle.py) against vectors recorded from the unmodified reference
(tools/make_golden.py).  CPU only.  This is what pins the oracle (task section 3)."""
import os

import pytest

import akshar_oracle as O


def test_pins_recorded(golden):
    assert golden['pins'] == {'regex': '2026.3.32', 'tokenizers': '0.22.2', 'sentencepiece': '0.2.1',
                              'unicodedata': '15.0.0'}
    assert O.tables().versions['regex'] == golden['pins']['regex']
    assert O.tables().versions['unicodedata'] == golden['pins']['unicodedata']


def test_normalize_family(golden):
    for r in golden['rows']:
        s = r['in']
        assert O.normalize_unicode(s) == r['nfc']
        assert O.semantic_normalize(s) == r['sem']
        assert O.filter_garbage(s) == r['filt']
        assert O.remove_elongations(s) == r['elong']
        assert O.normalize_text(s) == r['norm']
        assert O.normalize_text(s, normalize_roman=False) == r['norm_nr']
        assert O.normalize_text(s, clean_hinglish=False) == r['norm_nc']
        assert O.normalize_text(s, normalize_roman=False, clean_hinglish=False) == r['norm_raw']


def test_reference_test_expectations():
    # reference tests/test_normalize.py:31-54 and SURVEY.md section 0 gotcha 1
    assert O.semantic_normalize("Hello World") == "hello world"
    assert O.semantic_normalize("नमस्ते") == "नमस्ते"
    assert O.semantic_normalize("hello नमस्ते world") == "hello नमस्ते world"
    assert O.remove_elongations("yaaaaar") == "yar"
    assert O.remove_elongations("bohoooot") == "bohot"
    assert len(O.normalize_unicode("नमस्ते")) == len("नमस्ते")
    # reference tests/test_segment.py:37-54
    for ch, t in (('न', 'devanagari'), ('म', 'devanagari'), ('a', 'roman'), ('Z', 'roman'), ('5', 'digit'),
                  ('.', 'punct'), (' ', 'punct')):
        assert O.identify_script(ch) == t
    assert any('क्ष' in a for a in O.segment_akshars("क्षेत्रे"))
    # executed notebook cells
    assert O.segment_akshars(O.normalize_text("aaj मौसम बहुत अच्छा है")) == \
        ['a', 'a', 'j', ' ', 'मौ', 'स', 'म', ' ', 'ब', 'हु', 'त', ' ', 'अ', 'च्छा', ' ', 'है']
    assert [O.roman_phonetic_signature(w) for w in ('nahi', 'nahii', 'nahee')] == ['nahi', 'nahii', 'nahi']


def test_word_tokenizers(golden):
    # reference segment.py:239-401, recorded by tools/make_golden.py
    for r in golden['rows']:
        s = r['in']
        assert O.word_tokenize_hindi(s) == r['words_hi']
        assert O.word_tokenize_sanskrit(s) == r['words_sa']
        assert O.word_tokenize(s) == r['words_auto']
    # str.isspace() written out in the oracle == this interpreter's (Unicode 15.0, pinned above)
    assert sorted(O._PY_SPACE) == [c for c in range(0x110000) if chr(c).isspace()]
    assert O.word_tokenize("राम। सीता॥ (गीता)") == ['राम', '।', 'सीता', '॥', 'गीता']
    assert O.word_tokenize("a\u2003b\x1fc  d", language='xx') == ['a', 'b', 'c', 'd']


def test_feature_wrappers(golden):
    for r in golden['rows']:
        assert O.akshara_level_tokenization(r['in']) == r['feat_akshara']
        assert O.preserve_nukta(r['in']) == r['feat_nukta']


def test_segment_and_runs(golden):
    for r in golden['rows']:
        s, n = r['in'], r['norm']
        assert O.segment_akshars(s) == r['seg_raw']
        assert O.segment_akshars(n) == r['seg']
        assert O.segment_akshars(n, matras=True) == r['seg_m']
        assert O.segment_akshars(s, separate_matras=True) == r['segm_raw']
        assert [list(x) for x in O.detect_code_switches(s)] == r['cs_raw']
        assert [list(x) for x in O.detect_code_switches(n)] == r['cs']
        assert O.analyze_text_composition(n) == r['comp']
        assert O.segment_akshars(n) == r['tokenize']


def test_identify_script(golden):
    for ch, t in golden['identify_script'].items():
        assert O.identify_script(ch) == t


def test_signature(golden):
    for w, sig in golden['signature'].items():
        assert O.roman_phonetic_signature(w) == sig, w


@pytest.mark.parametrize('name,kind', [('bpe24k', 'bpe'), ('bpe_corpus', 'bpe'), ('spm24k', 'spm'), ('spm_corpus', 'spm')])
def test_subword_ids(golden, bpe_rows, models_dir, name, kind):
    if kind == 'bpe':
        m = O.BpeModel(os.path.join(models_dir, name + '.json'))
        enc = lambda t: O.bpe_encode(m, t)
    else:
        m = O.UnigramModel(os.path.join(models_dir, name + '.model'))
        enc = lambda t: O.unigram_encode(m, t)
    assert m.vocab_size() == golden['vocab_size'][name]
    rows = bpe_rows if kind == 'bpe' else golden['rows']
    assert len(rows) > 2000
    for r in rows:
        assert enc(r['norm']) == r['ids_' + name], r['in']
    if kind == 'bpe':
        # U+09FE (Unicode 11) is newer than HF's tables: no reordering around it there
        assert O.hf_nfkc_cps([0x61, 0x9FE, 0x323]) == [0x61, 0x9FE, 0x323] and O.nfc_cps([0x61, 0x9FE, 0x323]) == [0x1EA1, 0x9FE]


def test_pieces_and_decode(golden, bpe_rows, models_dir):
    bm = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    um = O.UnigramModel(os.path.join(models_dir, 'spm24k.model'))
    for r in golden['rows'][:600]:
        assert O.unigram_pieces(um, r['norm']) == r['pieces_spm24k']
    for r in bpe_rows[:600]:
        assert [bm.id_to_token[i] for i in r['ids_bpe24k']] == r['pieces_bpe24k']
        assert O.bpe_decode(bm, r['ids_bpe24k']) == r['dec_bpe24k']


def test_raw_mode_ids(golden_raw, models_dir):
    """clean_hinglish=False: text outside the closed alphabet reaches the models (emoji, accents, other scripts); on the BPE
    side HF's own NFKC and added-token matching act on it (scripts/train_bpe.py:71,80) -- every row, nothing filtered"""
    bm = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    um = O.UnigramModel(os.path.join(models_dir, 'spm24k.model'))
    for r in golden_raw['rows']:
        n = r['norm_nc']
        assert O.normalize_text(r['in'], clean_hinglish=False) == n
        assert O.unigram_encode(um, n) == r['ids_spm24k'], r['in']
        assert O.bpe_encode(bm, n) == r['ids_bpe24k'], r['in']
        raw = O.normalize_text(r['in'], normalize_roman=False, clean_hinglish=False)
        assert O.bpe_encode(bm, raw) == r['ids_bpe24k_raw'], r['in']
        assert ''.join(chr(c) for c in O.hf_nfkc_cps([ord(c) for c in r['in']])) == r['hf_nfkc'], r['in']


def test_decode_and_detokenize(golden, decode_fuzz, models_dir):
    """tokenizer.py:195-246: decode on the rows the encoders produce (golden) and on fuzzed id rows (decode_fuzz)"""
    bm = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    um = O.UnigramModel(os.path.join(models_dir, 'spm24k.model'))
    for r in golden['rows']:
        assert O.bpe_decode(bm, r['ids_bpe24k']) == r['dec_bpe24k']
        assert O.unigram_decode(um, r['ids_spm24k']) == r['dec_spm24k']
        assert O.bpe_detokenize(r['pieces_bpe24k']) == r['detok_bpe24k']
        assert O.spm_detokenize(r['pieces_spm24k']) == r['detok_spm24k']
    for name in ('spm24k', 'spm_corpus'):
        m = O.UnigramModel(os.path.join(models_dir, name + '.model'))
        F = decode_fuzz[name]
        for ids, dec, det in zip(F['ids'], F['decode'], F['detokenize']):
            assert O.unigram_decode(m, ids) == dec
            assert O.spm_detokenize([m.pieces[i][0] for i in ids]) == det
        with pytest.raises(IndexError):
            O.unigram_decode(m, [len(m.pieces)])
    for name in ('bpe24k', 'bpe_corpus'):
        m = O.BpeModel(os.path.join(models_dir, name + '.json'))
        F = decode_fuzz[name]
        for ids, dec in zip(F['ids'], F['decode']):
            assert O.bpe_decode(m, ids) == dec
        for ids, det in zip(F['ids_detok'], F['detokenize']):
            assert O.bpe_detokenize([m.id_to_token[i] for i in ids]) == det
    S = decode_fuzz['detokenize_strings']
    assert [O.bpe_detokenize(t) for t in S['tokens']] == S['bpe']
    assert [O.spm_detokenize(t) for t in S['tokens']] == S['sentencepiece']
