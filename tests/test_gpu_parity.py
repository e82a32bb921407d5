"""Parity of the CUDA path (through the C ABI) with (1) vectors recorded from the unmodified reference,
(2) the oracle on seeded synthetic / adversarial inputs, (3) size-independent properties at large sizes."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import akshar_oracle as O
import oracle_batch as OB
import synth_corpus as sc


@pytest.fixture(scope='module')
def A():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import __graft_entry__ as g
    g.build()
    import akshar_b200
    return akshar_b200


@pytest.fixture(scope='module')
def eng(A):
    return A.engine()


def _np(t):
    return t.cpu().numpy()


# ------------------------------------------------------------------ golden vectors (unmodified reference)
@pytest.mark.parametrize('nr,nc,key', [(True, True, 'norm'), (False, True, 'norm_nr'), (True, False, 'norm_nc'),
                                       (False, False, 'norm_raw')])
def test_normalize_golden(A, golden, nr, nc, key):
    ins = [r['in'] for r in golden['rows']]
    assert A.normalize_batch(ins, nr, nc) == [r[key] for r in golden['rows']]


@pytest.mark.parametrize('name,key', [('normalize_unicode', 'nfc'), ('semantic_normalize', 'sem'), ('filter_garbage', 'filt'),
                                      ('remove_elongations', 'elong')])
def test_single_stages_golden(A, golden, name, key):
    from akshar_b200.normalize import stage_batch
    ins = [r['in'] for r in golden['rows']]
    assert stage_batch(name, ins) == [r[key] for r in golden['rows']]


def test_segment_golden(A, golden):
    rows = golden['rows']
    norm = [r['norm'] for r in rows]
    raw = [r['in'] for r in rows]
    assert A.segment_akshars_batch(norm) == [r['seg'] for r in rows]
    assert A.segment_akshars_batch(raw) == [r['seg_raw'] for r in rows]
    assert A.segment_akshars_batch(norm, matras=True) == [r['seg_m'] for r in rows]
    assert A.segment_akshars_batch(raw, matras=True) == [r['segm_raw'] for r in rows]


def test_code_switch_golden(A, golden):
    rows = golden['rows']
    got = A.detect_code_switches_batch([r['norm'] for r in rows])
    assert [[list(x) for x in g] for g in got] == [r['cs'] for r in rows]
    got = A.detect_code_switches_batch([r['in'] for r in rows])
    assert [[list(x) for x in g] for g in got] == [r['cs_raw'] for r in rows]
    assert A.analyze_text_composition_batch([r['norm'] for r in rows]) == [r['comp'] for r in rows]


def test_word_tokenizers_golden(A, golden):
    """reference segment.py:239-401 (recorded from the unmodified reference): Hindi / Sanskrit loops and the auto route"""
    rows = golden['rows']
    ins = [r['in'] for r in rows]
    assert A.word_tokenize_hindi_batch(ins) == [r['words_hi'] for r in rows]
    assert A.word_tokenize_sanskrit_batch(ins) == [r['words_sa'] for r in rows]
    assert A.word_tokenize_batch(ins) == [r['words_auto'] for r in rows]
    assert A.word_tokenize_batch(ins, language='hi') == [r['words_hi'] for r in rows]
    assert A.word_tokenize_batch(ins[:500], language='tamil') == [s.split() for s in ins[:500]]
    assert A.word_tokenize("\u0930\u093e\u092e\u0964 \u0938\u0940\u0924\u093e\u0965 (x)") == \
        ['\u0930\u093e\u092e', '\u0964', '\u0938\u0940\u0924\u093e', '\u0965', 'x']
    assert A.word_tokenize("Hello,  World \u2003 again") == ['Hello,', 'World', 'again']
    assert A.word_tokenize("") == [] and A.word_tokenize_hindi("   ") == [] and A.word_tokenize_batch([]) == []


def test_word_tokenizers_oracle_large(A, eng):
    """64 MB of mixed text against the oracle's loops, both rules, row flags; then a property at 512 MB: every token is a
    maximal run of non-separator bytes (checked on the device output with numpy)"""
    lines = sc.Corpus('social', 31).lines(20 << 20) + sc.adversarial(3000, 77, 64) + sc.Corpus('hindi', 32).lines(20 << 20)
    from akshar_b200 import _lib as C
    b = eng.put(lines)
    import test_span_walkers as TW
    for mode, rule in ((0, C.WORDS_HINDI), (1, C.WORDS_SPLIT)):
        eb, ee, es, ef = TW._words_expected(lines, mode)
        wb, we, sp, fl = eng.word_tokenize_batch(b, rule=rule, row_flags=True)
        assert np.array_equal(_np(sp), es)
        assert np.array_equal(_np(wb), eb) and np.array_equal(_np(we), ee)
        assert np.array_equal(_np(fl), ef)
    big = sc.Corpus('hinglish', 33)
    data, off = big.generate(512 << 20)
    import torch
    tb = A.TextBatch(torch.from_numpy(data).cuda(), torch.from_numpy(off).cuda(), 0, int(off[-1]))
    wb, we, sp = eng.word_tokenize_batch(tb, rule=C.WORDS_SPLIT)
    spn = _np(sp)
    assert spn[0] == 0 and spn[-1] == wb.numel() and np.all(np.diff(spn) >= 0)
    # the corpus is ASCII + Devanagari (no wide spaces): words = non-space bytes after a space or a row start
    lut = np.zeros(256, dtype=bool)
    lut[[9, 10, 11, 12, 13, 28, 29, 30, 31, 32]] = True
    space = lut[data]
    starts = np.zeros(data.size + 1, dtype=bool)
    starts[off[:-1]] = True
    prev_space = np.concatenate(([True], space[:-1]))
    n_words = int(np.count_nonzero(~space & (prev_space | starts[:-1])))
    assert wb.numel() == n_words
    row_of = np.repeat(np.arange(off.size - 1), np.diff(spn))
    ab = off[row_of] + _np(wb)
    ae = off[row_of] + _np(we)
    assert np.all(ae > ab) and np.all(ab[1:] >= ae[:-1])
    assert int((ae - ab).sum()) == int(np.count_nonzero(~space))          # tokens cover exactly the non-space bytes


def test_composition_and_feature_wrappers_golden(A, golden):
    """analyze_text_composition (segment.py:210-236) from the fused per-row counts, tokenize(return_metadata) / explain values
    (tokenizer.py:123-165, 248-276), akshara_level_tokenization / preserve_nukta (features.py:28-55, 173-206)"""
    rows = golden['rows']
    norm = [r['norm'] for r in rows]
    ins = [r['in'] for r in rows]
    assert A.analyze_text_composition_batch(norm) == [r['comp'] for r in rows]
    assert A.akshara_level_tokenization_batch(ins) == [r['feat_akshara'] for r in rows]
    assert A.preserve_nukta_batch(ins) == [r['feat_nukta'] for r in rows]
    assert A.akshara_level_tokenization('') == [] and A.preserve_nukta('\u0915\u093c') == ['\u0915\u093c']
    assert A.analyze_text_composition('') == {'akshar_count': 0, 'script_switches': -1, 'devanagari_ratio': 0, 'roman_ratio': 0}
    fb = A.aksharTokenizer()
    for r in rows[:300]:
        assert fb.tokenize(r['in'], return_metadata=True) == r['meta']
        assert fb.detokenize(r['tokenize']) == r['detok_akshar']
    # tokenize without a model = normalize + akshars in one library call, boundaries back as bit masks
    assert fb.tokenize_batch([r['in'] for r in rows]) == [r['tokenize'] for r in rows]
    for r in rows[::37]:
        assert fb.tokenize(r['in']) == r['tokenize']
    assert fb.tokenize_batch([]) == [] and fb.tokenize('') == []
    import akshar.features as F
    assert F.preserve_nukta is A.preserve_nukta
    tb = A.aksharTokenizer(os.path.join(os.path.dirname(__file__), 'golden', 'models', 'bpe24k.json'), 'bpe')
    for r in rows[:150]:
        ex = tb.explain(r['in'])
        ex['code_switches'] = [list(x) for x in ex['code_switches']]
        assert ex == r['explain_bpe24k']
    exb = tb.explain_batch([r['in'] for r in rows[:600]])
    for ex, r in zip(exb, rows[:600]):
        ex['code_switches'] = [list(x) for x in ex['code_switches']]
        assert ex == r['explain_bpe24k']
    # adversarial: runs of nukta / halant clusters of every length and parity, across rows
    lines = ['', '\u0915\u093c' * 7, 'a' + '\u0915\u093c' * 4 + 'b', '\u0915\u094d ' * 5, '\u0915\u094d\u0937 \u0924\u094d \u0930\u094d', '',
             '\u0915\u093c', 'x'] + sc.adversarial(3000, 5, 40)
    assert A.akshara_level_tokenization_batch(lines) == [O.akshara_level_tokenization(s) for s in lines]
    assert A.preserve_nukta_batch(lines) == [O.preserve_nukta(s) for s in lines]
    assert A.analyze_text_composition_batch(lines) == [O.analyze_text_composition(s) for s in lines]


def test_identify_script_full_table(A, golden):
    """identify_script (segment.py:128-147) for the 12 292 characters recorded from the reference, on the device: one row per
    character through the script-run kernel"""
    chars = list(golden['identify_script'])
    runs = A.engine().segment_batch(chars, clusters=False, runs=True)[1]
    tags = _np(runs.extra)
    sp = _np(runs.splits)
    assert np.all(np.diff(sp) == 1)
    names = ['devanagari', 'roman', 'digit', 'punct', 'other']
    from akshar_b200.segment import _PUNCT
    for ch, t in zip(chars, tags.tolist()):
        got = names[t] if t != 255 else ('punct' if ch in _PUNCT else 'digit')
        assert got == golden['identify_script'][ch], hex(ord(ch))
    for ch in chars[::97]:
        assert A.identify_script(ch) == golden['identify_script'][ch]


def _mask_words(pos, n_bytes):
    m = np.zeros(((n_bytes + 32) // 32) * 32, dtype=bool)
    m[pos] = True
    return np.packbits(m, bitorder='little').view(np.uint32)


def test_segment_bit_masks(A, eng, golden):
    """AKSHAR_SEG_MASK: the cluster / run boundaries as one bit per text byte + two tag planes == the offset form of the same
    call, on golden rows (normalized and raw), adversarial text (slow lanes), empty rows and a text that does not start on a
    16-byte boundary"""
    import torch
    sets = [[r['norm'] for r in golden['rows']], [r['in'] for r in golden['rows']],
            ['', '', 'a', '', '\u0915\u094d\u0937', ''] + sc.adversarial(4000, 3, 60) + ['', ''], [''], ['', '', '']]
    for lines in sets:
        for matras in (False, True):
            data, off = sc.pack(lines)
            for lead in (0, 5):
                d = torch.from_numpy(np.concatenate([np.full(lead, 32, dtype=np.uint8), data])).cuda()
                o = torch.from_numpy(off + lead).cuda()
                tb = A.TextBatch(d, o, lead, lead + int(off[-1]))
                cl, ru = eng.segment_batch(tb, clusters=True, matras=matras, runs=True)
                mk = eng.segment_masks(tb, clusters=True, matras=matras, runs=True)
                n = int(off[-1])
                res = _np(mk['result'])
                assert res[2] == 0 and res[0] == cl.values.numel() and res[1] == ru.values.numel()
                for rag, key in ((cl, 'cluster'), (ru, 'run')):
                    sp = _np(rag.splits)
                    row_of = np.repeat(np.arange(len(sp) - 1), np.diff(sp))
                    pos = off[row_of] + _np(rag.values)
                    assert np.array_equal(_np(mk[key]).view(np.uint32), _mask_words(pos, n))
                tags = _np(ru.extra)
                t0 = _mask_words(pos[(tags == 1) | (tags == 255)], n)
                t1 = _mask_words(pos[(tags == 4) | (tags == 255)], n)
                planes = _np(mk['tags']).view(np.uint32)
                assert np.array_equal(planes[0], t0) and np.array_equal(planes[1], t1)


def test_pipeline_host_pipelined(A, eng, golden):
    """normalize -> akshars -> script runs through the pipelined host path (pinned text in; normalized text, row offsets and
    boundary bit masks out) == the reference vectors, and == the two-call offset form on 48 MB over many chunks"""
    import torch
    rows = golden['rows']
    data, off = sc.pack([r['in'] for r in rows])
    hd, ho = torch.from_numpy(data).pin_memory(), torch.from_numpy(off).pin_memory()
    out = eng.pipeline_host_pipelined(hd, ho, chunk_bytes=1 << 16)
    nb = out.norm.numpy().tobytes()
    no = out.offsets.numpy()
    assert [nb[no[i]:no[i + 1]].decode('utf-8') for i in range(len(rows))] == [r['norm'] for r in rows]
    ce, cs = out.cluster_ends()
    re_, rs, rt = out.run_ends()
    names = ['devanagari', 'roman', 'digit', 'punct', 'other']
    for i, r in enumerate(rows):
        b = r['norm'].encode('utf-8')
        ends = ce[cs[i]:cs[i + 1]].tolist()
        assert [b[(ends[j - 1] if j else 0):e].decode('utf-8') for j, e in enumerate(ends)] == r['seg']
        ends = re_[rs[i]:rs[i + 1]].tolist()
        labs = [None if t == 255 else names[t] for t in rt[rs[i]:rs[i + 1]].tolist()]
        assert [[b[(ends[j - 1] if j else 0):e].decode('utf-8'), labs[j]] for j, e in enumerate(ends)] == r['cs']
    assert out.n_clusters == ce.size and out.n_runs == re_.size
    # 48 MB of social text + adversarial rows (the row-sequential fallback of a chunk included), 2 MiB chunks
    lines = sc.Corpus('social', 51).lines(48 << 20) + sc.adversarial(3000, 8, 64) + ['', 'x']
    data, off = sc.pack(lines)
    hd, ho = torch.from_numpy(data).pin_memory(), torch.from_numpy(off).pin_memory()
    out = eng.pipeline_host_pipelined(hd, ho, chunk_bytes=2 << 20)
    norm = eng.normalize_batch((hd, ho))
    cl, ru = eng.segment_batch(norm, clusters=True, runs=True)
    assert out.norm.numel() == norm.end and np.array_equal(out.norm.numpy(), _np(norm.data[:norm.end]))
    assert np.array_equal(out.offsets.numpy(), _np(norm.offsets))
    ce, cs = out.cluster_ends()
    assert np.array_equal(cs, _np(cl.splits)) and np.array_equal(ce, _np(cl.values))
    re_, rs, rt = out.run_ends()
    assert np.array_equal(rs, _np(ru.splits)) and np.array_equal(re_, _np(ru.values)) and np.array_equal(rt, _np(ru.extra))


def test_signature_golden(A, golden):
    words = list(golden['signature'])
    assert A.roman_phonetic_signature_batch(words) == [golden['signature'][w] for w in words]


@pytest.mark.parametrize('name,kind', [('bpe24k', 'bpe'), ('bpe_corpus', 'bpe'), ('spm24k', 'sentencepiece'),
                                       ('spm_corpus', 'sentencepiece')])
def test_encode_golden(A, golden, bpe_rows, models_dir, name, kind):
    path = os.path.join(models_dir, name + ('.json' if kind == 'bpe' else '.model'))
    tk = A.aksharTokenizer(path, kind)
    assert tk.vocab_size() == golden['vocab_size'][name]
    rows = bpe_rows if kind == 'bpe' else golden['rows']
    # raw text through the fused normalize + encode pipeline (aksharTokenizer.encode over a batch)
    assert tk.encode_batch([r['in'] for r in rows]) == [r['ids_' + name] for r in rows]
    # already-normalized rows through the stand-alone encoders
    e = tk._eng.encode_bpe_batch if kind == 'bpe' else tk._eng.encode_unigram_batch
    got = [x.tolist() for x in e([r['norm'] for r in rows]).rows()]
    assert got == [r['ids_' + name] for r in rows]


def test_pieces_and_decode_golden(A, golden, bpe_rows, models_dir):
    tb = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe')
    tu = A.aksharTokenizer(os.path.join(models_dir, 'spm24k.model'))
    rows = bpe_rows[:400]
    assert tb.tokenize_batch([r['in'] for r in rows]) == [r['pieces_bpe24k'] for r in rows]
    rows = golden['rows'][:400]
    assert tu.tokenize_batch([r['in'] for r in rows]) == [r['pieces_spm24k'] for r in rows]
    rows = bpe_rows[:400]
    for r in rows[:200]:
        assert tb.decode(r['ids_bpe24k']) == r['dec_bpe24k']
        assert tu.decode(r['ids_spm24k']) == r['dec_spm24k']


@pytest.mark.parametrize('name,kind', [('spm24k', 'sentencepiece'), ('spm_corpus', 'sentencepiece'), ('bpe24k', 'bpe'), ('bpe_corpus', 'bpe')])
def test_decode_on_device(A, golden, decode_fuzz, bpe_rows, models_dir, name, kind):
    """tokenizer.py:195-246 on the device: what the unmodified reference decoded / detokenized, on the ids the encoders
    produce and on fuzzed id rows (byte pieces out of order, control / unknown / special / unused ids)"""
    tk = A.aksharTokenizer(os.path.join(models_dir, name + ('.json' if kind == 'bpe' else '.model')), kind)
    F = decode_fuzz[name]
    assert tk.decode_batch(F['ids']) == F['decode']
    assert tk.detokenize_batch(F['ids_detok'] if kind == 'bpe' else F['ids']) == F['detokenize']
    for ids, dec in list(zip(F['ids'], F['decode']))[:40]:
        assert tk.decode(ids) == dec
    if name in ('spm24k', 'bpe24k'):
        rows = bpe_rows if kind == 'bpe' else golden['rows']
        assert tk.decode_batch([r['ids_' + name] for r in rows]) == [r['dec_' + name] for r in rows]
        assert tk.detokenize_batch([r['ids_' + name] for r in rows]) == [r['detok_' + name] for r in rows]
        for r in rows[:60]:
            assert tk.detokenize(r['pieces_' + name]) == r['detok_' + name]
        S = decode_fuzz['detokenize_strings']
        assert [tk.detokenize(t) for t in S['tokens']] == S[kind]
        # straight from the encoder's device output, int32 and compact uint16 ids
        ins = [r['in'] for r in rows[:3000]]
        ids, _ = tk.encode_batch(ins, as_device=True)
        assert tk.decode_batch(ids) == [r['dec_' + name] for r in rows[:3000]]
        import torch
        assert tk.decode_batch((ids.values.to(torch.uint16), ids.splits)) == [r['dec_' + name] for r in rows[:3000]]
    if kind == 'sentencepiece':
        with pytest.raises(IndexError):
            tk.decode([5, tk.vocab_size()])
    assert tk.decode_batch([]) == [] and tk.decode([]) == '' and tk.decode_batch([[], []]) == ['', '']
    assert A.aksharTokenizer().detokenize(['a', 'b']) == 'ab'
    with pytest.raises(ValueError):
        A.aksharTokenizer().decode([1])


def test_decode_round_trip_1gib(A, models_dir):
    """encode -> decode at 1 GiB: for text that normalize_text has produced (single spaces, nothing to strip) the Unigram
    model's DecodeIds(EncodeAsIds(x)) is x, byte for byte -- compared on the device"""
    import torch
    tk = A.aksharTokenizer(os.path.join(models_dir, 'spm24k.model'), 'sentencepiece')
    eng = tk._eng
    data, off = sc.Corpus('hindi', 41).generate(1 << 30)
    raw = A.TextBatch(torch.from_numpy(data).cuda(), torch.from_numpy(off).cuda(), 0, int(off[-1]))
    norm = eng.normalize_batch(raw)
    del raw
    ids = eng.encode_unigram_batch(norm)
    eng.timing(True)
    back = eng.decode_batch(ids, None, 1)
    ms = eng.kernel_ms('ak_dec_kernel')
    eng.timing(False)
    assert back.end == norm.end
    assert torch.equal(back.offsets, norm.offsets)
    assert torch.equal(back.data[:back.end], norm.data[:norm.end])
    print('decode 1 GiB: %d ids, write kernel %.2f ms' % (ids.values.numel(), ms))


# ------------------------------------------------------------------ the reference's own test expectations, via the drop-in API
def test_reference_unit_expectations(A):
    # reference tests/test_normalize.py
    assert len(A.normalize_unicode("नमस्ते")) == len("नमस्ते")
    assert A.semantic_normalize("Hello World") == "hello world"
    assert A.semantic_normalize("नमस्ते") == "नमस्ते"
    assert A.semantic_normalize("hello नमस्ते world") == "hello नमस्ते world"
    for s in ("yaaaaar", "bohoooot", "hiiii", "normal"):
        assert len(A.remove_elongations(s)) <= len(s)
    assert A.remove_elongations("yaaaaar") == "yar"
    assert isinstance(A.roman_phonetic_signature("nahi"), str)
    r = A.normalize_text("Heyyy यार kya HAAL hai")
    assert "यार" in r and "heyyy" not in r
    # reference tests/test_segment.py
    assert A.segment_akshars("") == []
    assert len(A.segment_akshars("नमस्ते")) > 0
    assert any('क्ष' in a for a in A.segment_akshars("क्षेत्रे"))
    for ch, t in (('न', 'devanagari'), ('म', 'devanagari'), ('a', 'roman'), ('Z', 'roman'), ('5', 'digit'), ('.', 'punct'),
                  (' ', 'punct')):
        assert A.identify_script(ch) == t
    segs = A.detect_code_switches("aaj मौसम अच्छा hai")
    assert len(segs) > 1 and {s for _, s in segs} >= {'roman', 'devanagari'}
    assert A.detect_code_switches("") == []
    assert A.segment_by_script("hello नमस्ते world")
    st = A.analyze_text_composition("aaj मौसम अच्छा hai")
    assert set(st) >= {'akshar_count', 'script_switches', 'devanagari_ratio', 'roman_ratio'}
    # reference tests/test_tokenizer.py
    tk = A.AksharTokenizer()
    assert tk.model is None and tk.vocab_size() == 0
    assert isinstance(tk.tokenize("नमस्ते"), list)
    meta = tk.tokenize("hello नमस्ते", return_metadata=True)
    assert 'tokens' in meta and meta['token_count'] == len(meta['tokens'])
    ex = tk.explain("aaj मौसम अच्छा hai")
    assert set(ex) == {'original', 'normalized', 'akshars', 'code_switches', 'tokens', 'stats'}
    with pytest.raises(ValueError):
        tk.encode("hello")
    with pytest.raises(ValueError):
        tk.decode([1, 2])
    assert tk.detokenize(['a', 'b']) == 'ab'
    with pytest.raises(ValueError):
        A.aksharTokenizer(os.path.join(os.path.dirname(__file__), 'conftest.py'), 'nonsense')
    # executed notebook cell
    assert tk.tokenize("aaj मौसम बहुत अच्छा है") == ['a', 'a', 'j', ' ', 'मौ', 'स', 'म', ' ', 'ब', 'हु', 'त', ' ', 'अ', 'च्छा', ' ', 'है']


def test_reference_test_files_run_unmodified(A):
    """the reference's three test files (byte-for-byte copies kept as fixtures under tests/golden/ref_tests) against the
    `akshar` drop-in package: `from akshar.tokenizer import AksharTokenizer`, `from akshar.normalize import ...`"""
    import unittest
    here = os.path.join(os.path.dirname(__file__), 'golden', 'ref_tests')
    import akshar
    assert akshar.aksharTokenizer is A.aksharTokenizer
    suite = unittest.defaultTestLoader.discover(here, pattern='test_*.py', top_level_dir=here)
    res = unittest.TextTestRunner(verbosity=0).run(suite)
    assert res.testsRun >= 25 and res.wasSuccessful(), (res.errors, res.failures)


# ------------------------------------------------------------------ oracle on seeded synthetic + adversarial input
def _mixed_lines():
    lines = sc.adversarial(3000, 31, 80) + sc.Corpus('social', 11).lines(300000) + sc.Corpus('hindi', 12).lines(200000)
    lines += ['', '', '\n\n\n', ' ', 'a' * 5000, 'क' + 'ु' * 300 + 'ख', '🇮' * 201, 'अ' + '्क' * 400, '']
    return lines


def test_oracle_parity_mixed(A, eng):
    lines = _mixed_lines()
    exp, exp_off = OB.normalize_batch(lines)
    tb = eng.normalize_batch(lines)
    assert np.array_equal(_np(tb.offsets), exp_off)
    assert _np(tb.data)[:tb.end].tobytes() == exp.tobytes()
    norm = tb.to_strings()
    for src in (lines, norm):
        c, r = eng.segment_batch(src, clusters=True, runs=True)
        ce, cs = OB.segment_batch(src)
        re_, rt, rs = OB.runs_batch(src)
        assert np.array_equal(_np(c.splits), cs) and np.array_equal(_np(c.values), ce)
        assert np.array_equal(_np(r.splits), rs) and np.array_equal(_np(r.values), re_) and np.array_equal(_np(r.extra), rt)
        m, _ = eng.segment_batch(src, clusters=True, matras=True)
        me, ms = OB.segment_batch(src, matras=True)
        assert np.array_equal(_np(m.splits), ms) and np.array_equal(_np(m.values), me)


def test_modes_agree(A, eng):
    from akshar_b200 import _lib as C
    lines = _mixed_lines()[:2500] + ['', 'x']
    a = eng.normalize_batch(lines, mode=C.MODE_TILES)
    b = eng.normalize_batch(lines, mode=C.MODE_ROWS)
    assert a.end == b.end and np.array_equal(_np(a.offsets), _np(b.offsets))
    assert np.array_equal(_np(a.data)[:a.end], _np(b.data)[:b.end])
    ca, ra = eng.segment_batch(lines, runs=True, mode=C.MODE_TILES)
    cb, rb = eng.segment_batch(lines, runs=True, mode=C.MODE_ROWS)
    assert np.array_equal(_np(ca.values), _np(cb.values)) and np.array_equal(_np(ra.values), _np(rb.values))


def test_long_combining_run_falls_back_to_rows(A, eng):
    # > AK_LOOKBACK_LIMIT bytes of Extend characters: the tile path gives up loudly and the engine re-runs row-wise
    s = 'क' + 'ु' * 3000 + 'ख'
    assert A.segment_akshars_batch([s, 'ab']) == [O.segment_akshars(s), ['a', 'b']]
    # the mask form of the same text: through the offset form's row mode, turned into masks
    mk = eng.segment_masks([s, 'ab'], clusters=True, runs=True)
    cl, ru = eng.segment_batch([s, 'ab'], clusters=True, runs=True)
    n = len(s.encode('utf-8')) + 2
    off = np.array([0, n - 2, n])
    for rag, key in ((cl, 'cluster'), (ru, 'run')):
        sp = _np(rag.splits)
        pos = off[np.repeat(np.arange(2), np.diff(sp))] + _np(rag.values)
        assert np.array_equal(_np(mk[key]).view(np.uint32), _mask_words(pos, n))


def test_edge_batches(A, eng, models_dir):
    assert A.normalize_batch([]) == []
    assert A.segment_akshars_batch([]) == []
    assert A.normalize_batch(['', '', '']) == ['', '', '']
    assert A.segment_akshars_batch(['', 'a', '']) == [[], ['a'], []]
    assert A.detect_code_switches_batch(['', '123', '']) == [[], [('123', None)], []]
    tk = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe')
    assert tk.encode_batch(['', '']) == [[2, 3], [2, 3]]
    assert tk.encode_batch([]) == []
    tu = A.aksharTokenizer(os.path.join(models_dir, 'spm24k.model'))
    assert tu.encode_batch(['', ' ', 'a']) [:2] == [[], []]
    # capacities far too small: totals stay exact and the engine retries
    lines = sc.Corpus('hinglish', 2).lines(50000)
    tb = eng.normalize_batch(lines, capacity=16)
    assert tb.to_strings() == [O.normalize_text(s) for s in lines]
    c, _ = eng.segment_batch(lines, capacity=3)
    assert np.array_equal(_np(c.values), OB.segment_batch(lines)[0])
    r = tk._eng.encode_bpe_batch(tb.to_strings(), capacity=5)
    om = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    assert [x.tolist() for x in r.rows()] == [O.bpe_encode(om, s) for s in tb.to_strings()]


def test_pipelined_host_encode(A, models_dir):
    import torch
    tk = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe')
    data, off = sc.Corpus('social', 21).generate(24 << 20)
    h_data, h_off = torch.from_numpy(data).pin_memory(), torch.from_numpy(off).pin_memory()
    ids, splits = tk._eng.encode_host_pipelined(h_data, h_off, 0, chunk_bytes=3 << 20)       # 8+ chunks
    ref, _ = tk._eng.tokenizer_encode_batch((h_data, h_off), 0)
    assert torch.equal(ids, ref.values.cpu()) and torch.equal(splits, ref.splits.cpu())
    tu = A.aksharTokenizer(os.path.join(models_dir, 'spm24k.model'))
    ids, splits = tu.encode_batch_host(h_data, h_off)
    ref, _ = tu._eng.tokenizer_encode_batch((h_data, h_off), 1)
    assert torch.equal(ids, ref.values.cpu()) and torch.equal(splits, ref.splits.cpu())


def test_bpe_dense_events(A, models_dir):
    # more row / word starts in one 480-byte warp tile than its event list holds, and tens of thousands of empty
    # rows at one byte position (16-bit token offsets): both take the lane-by-lane path inside the fast kernel
    tk = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe')
    om = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    lines = ['a'] * 3000 + ['b.c,d!e'] * 200 + [''] * 40000 + ['kya haal hai', '']
    got = [x.tolist() for x in tk._eng.encode_bpe_batch(lines).rows()]
    assert got == [O.bpe_encode(om, s) for s in lines]
    assert tk.encode_batch(lines) == got


def test_bpe_long_words_every_alignment(A, models_dir):
    # word ends beyond the boundary masks a lane sees (its 32 bytes + the next two lanes'), at every alignment
    import numpy as np
    tk = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe')
    om = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    rng = np.random.default_rng(11)
    lines = []
    for _ in range(600):
        parts = []
        for _ in range(int(rng.integers(1, 6))):
            n = int(rng.integers(40, 120))
            parts.append(('ab' * 80)[:n] if rng.random() < 0.5 else ('\u0915\u0916' * 40)[:n // 3])
            parts.append('.' if rng.random() < 0.5 else ' ')
            parts.append('c' * int(rng.integers(0, 3)))
        lines.append(''.join(parts))
    for s0 in range(896, 928, 3):
        for end in (988, 989, 990, 991, 992):
            lines.append('y' * (s0 - 2) + '. ' + 'a' * (end - s0) + '.' + 'c' * 20)
    got = [x.tolist() for x in tk._eng.encode_bpe_batch(lines).rows()]
    assert got == [O.bpe_encode(om, s) for s in lines]
    for s in lines[-55:]:           # each crafted row as its own batch: it then starts at byte 0 of the first warp
        assert tk._eng.encode_bpe_batch([s]).rows()[0].tolist() == O.bpe_encode(om, s)


def test_bpe_renormalizes_on_device(A, models_dir):
    # rows that are not in NFC make the kernel take its conditional NFC + re-encode passes
    tk = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe', clean_hinglish=True)
    om = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    lines = ['ऩ क्ष', 'abc', 'ऱा', 'x'] * 50
    got = [x.tolist() for x in tk._eng.encode_bpe_batch(lines).rows()]
    assert got == [O.bpe_encode(om, s) for s in lines]
    long_word = ['ab' * 200 + ' ' + 'कख' * 90, 'ok']
    got = [x.tolist() for x in tk._eng.encode_bpe_batch(long_word).rows()]
    assert got == [O.bpe_encode(om, s) for s in long_word]


# ------------------------------------------------------------------ large sizes: size-independent properties
def test_large_properties(A, eng, models_dir):
    import torch
    nbytes = 256 << 20
    data, off = sc.Corpus('social', 77).generate(nbytes)
    host = (torch.from_numpy(data), torch.from_numpy(off))
    tb_in = eng.put(host)
    norm = eng.normalize_batch(tb_in)
    # (1) idempotence: normalize_text(normalize_text(x)) == normalize_text(x)
    again = eng.normalize_batch(norm)
    assert again.end == norm.end
    assert torch.equal(again.data[:again.end], norm.data[:norm.end]) and torch.equal(again.offsets, norm.offsets)
    # (2) batch-split invariance: the two halves processed separately give the same bytes / clusters / ids
    n = off.size - 1
    h = n // 2
    for lo, hi in ((0, h), (h, n)):
        sub = (torch.from_numpy(data[off[lo]:off[hi]].copy()), torch.from_numpy(off[lo:hi + 1] - off[lo]))
        part = eng.normalize_batch(sub)
        o = norm.offsets[lo:hi + 1]
        assert torch.equal(part.offsets, o - o[0])
        assert torch.equal(part.data[:part.end], norm.data[int(o[0]):int(o[-1])])
    # (3) cluster ends are strictly increasing inside a row and end at the row length
    c, r = eng.segment_batch(norm, clusters=True, runs=True)
    lens = (norm.offsets[1:] - norm.offsets[:-1])
    last = c.values[(c.splits[1:] - 1).clamp(min=0)]
    nonempty = lens > 0
    assert torch.equal(last[nonempty].long(), lens[nonempty])
    d = c.values[1:] - c.values[:-1]
    is_first = torch.zeros(c.values.numel(), dtype=torch.bool, device=c.values.device)
    is_first[c.splits[:-1][nonempty]] = True
    assert bool((d[~is_first[1:]] > 0).all())
    # (4) oracle on a seeded sample of rows
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(n, size=3000, replace=False))
    b = data.tobytes()
    nb = norm.data[:norm.end].cpu().numpy().tobytes()
    no = norm.offsets.cpu().numpy()
    ce = c.values.cpu().numpy()
    cs = c.splits.cpu().numpy()
    tk = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe')
    ids, _ = tk._eng.tokenizer_encode_batch(tb_in, 0)
    iv, isp = ids.values.cpu().numpy(), ids.splits.cpu().numpy()
    om = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    for i in pick:
        s = b[off[i]:off[i + 1]].decode('utf-8')
        e = O.normalize_text(s)
        assert nb[no[i]:no[i + 1]].decode('utf-8') == e
        cps = [ord(ch) for ch in e]
        assert ce[cs[i]:cs[i + 1]].tolist() == O.cp_ends_to_byte_ends(cps, O.segment_breaks(cps))
        assert iv[isp[i]:isp[i + 1]].tolist() == O.bpe_encode(om, e)


def test_unigram_at_scale_and_long_rows(A, models_dir):
    """the Unigram encoder beyond the golden rows: (1) 64 MB of in-vocabulary Hindi, every row against the oracle's Viterbi
    (sampled 6000 rows) and both modes against each other on all rows; (2) one 8 MB row (a file as one string: the exact
    whole-row Viterbi) and a batch with 40 kB rows between short ones; (3) compact outputs (uint16 ids, int32 splits)"""
    import torch
    tk = A.aksharTokenizer(os.path.join(models_dir, 'spm24k.model'), 'sentencepiece')
    eng = tk._eng
    um = O.UnigramModel(os.path.join(models_dir, 'spm24k.model'))
    from akshar_b200 import _lib as C
    data, off = sc.Corpus('hindi', 61).generate(64 << 20)
    host = (torch.from_numpy(data), torch.from_numpy(off))
    ids, norm = eng.tokenizer_encode_batch(host, 1)
    iv, isp = _np(ids.values), _np(ids.splits)
    rows_ids, _ = eng.tokenizer_encode_batch(host, 1, mode=C.MODE_ROWS)
    assert np.array_equal(_np(rows_ids.splits), isp) and np.array_equal(_np(rows_ids.values), iv)
    n = off.size - 1
    b = data.tobytes()
    rng = np.random.default_rng(7)
    for i in np.sort(rng.choice(n, size=6000, replace=False)):
        e = O.normalize_text(b[off[i]:off[i + 1]].decode('utf-8'))
        assert iv[isp[i]:isp[i + 1]].tolist() == O.unigram_encode(um, e), i
    # (2) long rows: rows beyond 8192 bytes take the exact row encoder
    lines = sc.Corpus('hindi', 62).lines(9 << 20)
    one = ' '.join(lines)[:8 << 20]
    mid = [' '.join(lines[k:k + 400]) for k in range(0, 4000, 400)]
    batch = lines[:50] + [one] + lines[50:100] + mid + ['', 'x']
    got = tk.encode_batch(batch)
    exp_short = [O.unigram_encode(um, O.normalize_text(s)) for s in lines[:100] + mid + ['', 'x']]
    assert got[:50] + got[51:] == exp_short
    # the 8 MB row: against the row-sequential mode, and against the oracle on its first 200 kB (a lattice cut at a space)
    alone, _ = eng.tokenizer_encode_batch([one], 1, mode=C.MODE_ROWS)
    assert got[50] == _np(alone.values).tolist()
    head = one[:200000]
    head = head[:head.rindex(' ')]
    eh = O.unigram_encode(um, O.normalize_text(head))
    assert got[50][:len(eh)] == eh
    assert tk.decode(got[50]) == O.normalize_text(one)
    # (3) compact outputs through the pipelined host path == the wide ones
    h_data, h_off = torch.from_numpy(data[:off[200000]]).pin_memory(), torch.from_numpy(off[:200001]).pin_memory()
    wide_ids, wide_sp = eng.encode_host_pipelined(h_data, h_off, 1, chunk_bytes=4 << 20)
    wide_ids, wide_sp = wide_ids.clone(), wide_sp.clone()
    cp = eng.encode_host_pipelined(h_data, h_off, 1, chunk_bytes=4 << 20, compact=True)
    assert np.array_equal(cp.ids().astype(np.int32), wide_ids.numpy()) and np.array_equal(cp.row_splits(), wide_sp.numpy())
    assert np.array_equal(wide_ids.numpy(), iv[:isp[200000]]) and np.array_equal(wide_sp.numpy(), isp[:200001])


def test_raw_mode_golden(A, golden_raw, models_dir):
    """clean_hinglish=False (reference vectors, every row): emoji, accents and other scripts reach the models, and on the
    BPE side what only HF acts on -- added tokens in the raw text, NFKC (scripts/train_bpe.py:71,80) -- is done on the device"""
    rows = golden_raw['rows']
    ins = [r['in'] for r in rows]
    tu = A.aksharTokenizer(os.path.join(models_dir, 'spm24k.model'), 'sentencepiece', clean_hinglish=False)
    assert tu.encode_batch(ins) == [r['ids_spm24k'] for r in rows]
    tb = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe', clean_hinglish=False)
    assert tb.encode_batch(ins) == [r['ids_bpe24k'] for r in rows]
    tb2 = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe', normalize_roman=False, clean_hinglish=False)
    assert tb2.encode_batch(ins) == [r['ids_bpe24k_raw'] for r in rows]
    for s in ('<s> hi </s>', 'x\ufb01y <mask>', 'a<pad>b<unk>c<', '\uff26\uff55\uff4c\uff4c \u2460 x\u00b2'):
        om = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
        assert tb2.encode(s) == O.bpe_encode(om, O.normalize_text(s, False, False))
    # the stand-alone encoder over text as it is (no normalize_text before it), in both modes
    from akshar_b200 import _lib as C
    om = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    exp = [O.bpe_encode(om, s) for s in ins]
    for mode in (C.MODE_TILES, C.MODE_ROWS):
        got = [x.tolist() for x in tb._eng.encode_bpe_batch(ins, mode=mode).rows()]
        assert got == exp


def test_corpus_front_end(A, models_dir, tmp_path):
    """the file front end (reference cli.py:46-84, 165-190): lines are found, stripped and dropped when empty ON THE DEVICE
    from the file's bytes (akshar_lines_batch); the file is read in pieces into pinned memory, no Python per line"""
    from akshar_b200 import corpus
    lines = sc.Corpus('social', 31).lines(400000)
    src = tmp_path / 'corpus.txt'
    seps = ['\n', '\r\n', '\r', '\n\n', '\n \t\n', '\n\u3000\n']
    body = ''.join((('  ' + l + ' \t') if i % 7 == 0 else ('\u00a0' + l + '\u2003') if i % 11 == 0 else l) + seps[i % len(seps)]
                   for i, l in enumerate(lines))
    src.write_bytes((body + '\n\n  \n' + 'last line without newline').encode('utf-8'))
    with open(src, 'r', encoding='utf-8') as f:
        ref_rows = [ln.strip() for ln in f.readlines() if ln.strip()]
    assert corpus.read_rows(str(src)) == ref_rows
    # pieces of 64 KiB: rows never straddle pieces; a single 300 KB line makes the reader grow its buffers
    big = tmp_path / 'big.txt'
    big.write_bytes(('x' * 300000 + '\n' + body).encode('utf-8'))
    got = []
    for rows in corpus.stream_rows(str(big), chunk_bytes=1 << 16):
        got.extend(rows.to_strings())
    assert got == ['x' * 300000] + ref_rows[:-1]
    out = tmp_path / 'corpus.preprocessed.txt'
    corpus.preprocess_corpus(str(src), str(out), chunk_bytes=1 << 17)
    exp = [O.normalize_text(l) for l in ref_rows]
    assert out.read_bytes().decode('utf-8') == ''.join(e + '\n' for e in exp)
    empty = tmp_path / 'empty.txt'
    empty.write_bytes(b'')
    assert corpus.read_rows(str(empty)) == []
    corpus.preprocess_corpus(str(empty), str(out))
    assert out.read_bytes() == b''
    # the whole file as ONE string, like `akshar tokenize -i FILE --format id` (one long row for the kernels)
    tk = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe')
    om = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    with open(src, 'r', encoding='utf-8') as f:
        text = f.read()
    assert corpus.tokenize_file(tk, str(src), 'id') == ' '.join(map(str, O.bpe_encode(om, O.normalize_text(text))))
    fb = A.aksharTokenizer()
    assert corpus.tokenize_file(fb, str(src), 'text') == ' '.join(O.segment_akshars(O.normalize_text(text)))
    enc = corpus.encode_lines(tk, str(src), chunk_bytes=1 << 17)
    assert len(enc) == len(exp) and enc[:50] == [O.bpe_encode(om, e) for e in exp[:50]] and enc[-1] == O.bpe_encode(om, exp[-1])


def test_full_size_bpe_1gib(A, models_dir):
    """BASELINE.json configs[1] at its full size (1 GiB synthetic Hinglish, BPE-24k): properties that do not need the
    oracle on every row -- per-row framing by <s> ... </s>, batch-split invariance through the pipelined host path,
    checksums of checksums -- plus the oracle on a seeded sample of rows."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
    import bench
    data, off = bench.make_corpus('hinglish', 1 << 30, 0)
    n = off.size - 1
    tk = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe')
    h_data, h_off = torch.from_numpy(data).pin_memory(), torch.from_numpy(off).pin_memory()
    ids, norm = tk._eng.tokenizer_encode_batch((h_data, h_off), 0)
    v, sp = ids.values, ids.splits
    assert int(sp[-1]) == v.numel() and int(sp[0]) == 0
    # every row is framed: first id <s> = 2, last id </s> = 3, and neither occurs anywhere else
    assert bool((v[sp[:-1]] == 2).all()) and bool((v[sp[1:] - 1] == 3).all())
    assert int((v == 2).sum()) == n and int((v == 3).sum()) == n
    # the same batch in 11 chunks through the pipelined host path: identical stream (a checksum of checksums would do;
    # the tensors are compared whole)
    pv, ps = tk._eng.encode_host_pipelined(h_data, h_off, 0, chunk_bytes=100 << 20)
    assert torch.equal(pv, v.cpu()) and torch.equal(ps, sp.cpu())
    # oracle on a seeded sample
    om = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    rng = np.random.default_rng(11)
    b = data
    hv, hs = pv.numpy(), ps.numpy()
    nb = norm.data[:norm.end].cpu().numpy()
    no = norm.offsets.cpu().numpy()
    for i in np.sort(rng.choice(n, size=2000, replace=False)):
        s = b[off[i]:off[i + 1]].tobytes().decode('utf-8')
        e = O.normalize_text(s)
        assert nb[no[i]:no[i + 1]].tobytes().decode('utf-8') == e
        assert hv[hs[i]:hs[i + 1]].tolist() == O.bpe_encode(om, e)


def test_cabi_error_codes(A, eng, models_dir):
    """status codes of the C ABI: bad arguments, workspace too small, encode before load -- nothing is enqueued"""
    import ctypes
    import torch
    from akshar_b200 import _lib as C
    from akshar_b200.batch import Engine
    lib = eng.lib
    b = eng.put(['hello', 'नमस्ते'])
    dev = eng.device
    out = torch.empty(64, dtype=torch.uint8, device=dev)
    off = torch.empty(3, dtype=torch.int64, device=dev)
    res = torch.empty(4, dtype=torch.int64, device=dev)
    ws = torch.empty(lib.akshar_workspace_bytes(b.n_bytes, 2), dtype=torch.uint8, device=dev)
    call = lambda flags, mode, wsz, text_end: lib.akshar_normalize_batch(
        eng._h, b.data.data_ptr(), b.offsets.data_ptr(), 2, 0, text_end, flags, mode, out.data_ptr(), 64, off.data_ptr(),
        res.data_ptr(), ws.data_ptr(), wsz, None)
    assert call(7, 0, ws.numel(), b.end) == C.OK
    assert call(1 << 6, 0, ws.numel(), b.end) == C.E_ARG            # unknown flag bit
    assert call(7, 5, ws.numel(), b.end) == C.E_ARG                 # unknown mode
    assert call(7, 0, ws.numel(), -1) == C.E_ARG                    # text_end < text_begin
    assert call(7, 0, 16, b.end) == C.E_WORKSPACE
    assert call(7, 0, ws.numel(), 1 << 33) == C.E_ARG               # more than one call takes (4 GiB - 64 KiB)
    assert b'too large' in lib.akshar_last_error(eng._h)
    assert call(7, 0, 16, b.end) == C.E_WORKSPACE
    assert b'workspace' in lib.akshar_last_error(eng._h)
    fresh = Engine(0)                                               # a context without models
    ids = torch.empty(64, dtype=torch.int32, device=dev)
    rc = lib.akshar_encode_bpe_batch(fresh._h, b.data.data_ptr(), b.offsets.data_ptr(), 2, 0, b.end, 0, ids.data_ptr(), 64,
                                     off.data_ptr(), res.data_ptr(), ws.data_ptr(), ws.numel(), None)
    assert rc == C.E_NOMODEL
    assert lib.akshar_vocab_size(fresh._h, 0) == C.E_NOMODEL and lib.akshar_vocab_size(fresh._h, 7) == C.E_ARG
    assert lib.akshar_load_bpe_json(fresh._h, b'{"not": "a tokenizer"}', 22) == C.E_MODEL
    assert lib.akshar_load_spm_model(fresh._h, b'\x00\x01\x02', 3) == C.E_MODEL
    # the round-2 entry points: same conventions
    i32 = torch.empty(64, dtype=torch.int32, device=dev)
    sp = torch.zeros(3, dtype=torch.int64, device=dev)
    assert lib.akshar_word_tokenize_batch(eng._h, b.data.data_ptr(), b.offsets.data_ptr(), 2, 0, b.end, 9, i32.data_ptr(),
                                          i32.data_ptr(), 64, sp.data_ptr(), None, res.data_ptr(), ws.data_ptr(), ws.numel(), None) == C.E_ARG
    assert lib.akshar_word_tokenize_batch(eng._h, b.data.data_ptr(), b.offsets.data_ptr(), 2, 0, b.end, 0, i32.data_ptr(),
                                          i32.data_ptr(), 64, sp.data_ptr(), None, res.data_ptr(), ws.data_ptr(), 16, None) == C.E_WORKSPACE
    assert lib.akshar_decode_batch(fresh._h, 0, 0, i32.data_ptr(), 0, 2, sp.data_ptr(), 2, out.data_ptr(), 64, off.data_ptr(),
                                   res.data_ptr(), ws.data_ptr(), ws.numel(), None) == C.E_NOMODEL
    assert lib.akshar_decode_batch(eng._h, 3, 0, i32.data_ptr(), 0, 2, sp.data_ptr(), 2, out.data_ptr(), 64, off.data_ptr(),
                                   res.data_ptr(), ws.data_ptr(), ws.numel(), None) == C.E_ARG
    assert lib.akshar_lines_batch(eng._h, b.data.data_ptr(), b.end, out.data_ptr(), 64, off.data_ptr(), 2, res.data_ptr(),
                                  ws.data_ptr(), 8, None) == C.E_WORKSPACE
    assert lib.akshar_lines_batch(eng._h, None, 5, out.data_ptr(), 64, off.data_ptr(), 2, res.data_ptr(), ws.data_ptr(),
                                  ws.numel(), None) == C.E_ARG
    assert lib.akshar_segment_batch(eng._h, b.data.data_ptr(), b.offsets.data_ptr(), 2, 0, b.end, C.SEG_CLUSTERS | C.SEG_MASK,
                                    C.MODE_ROWS, i32.data_ptr(), 64, None, None, None, 0, None, res.data_ptr(), ws.data_ptr(),
                                    ws.numel(), None) == C.E_ARG           # masks: tile mode only
    assert lib.akshar_segment_batch(eng._h, b.data.data_ptr(), b.offsets.data_ptr(), 2, 0, b.end, C.SEG_CLUSTERS | C.SEG_MASK,
                                    C.MODE_TILES, i32.data_ptr(), 0, None, None, None, 0, None, res.data_ptr(), ws.data_ptr(),
                                    ws.numel(), None) == C.E_ARG           # mask capacity in words too small
    assert lib.akshar_merge_clusters_batch(eng._h, b.data.data_ptr(), b.offsets.data_ptr(), 2, i32.data_ptr(), sp.data_ptr(), 0, 5,
                                           i32.data_ptr(), 64, sp.data_ptr(), res.data_ptr(), ws.data_ptr(), ws.numel(), None) == C.E_ARG
    assert lib.akshar_join_rows(eng._h, b.data.data_ptr(), b.offsets.data_ptr(), 2, 300, out.data_ptr(), None) == C.E_ARG
    # a row too short for the output capacity: totals exact, AKSHAR_ST_OVERFLOW, nothing written out of bounds
    assert lib.akshar_word_tokenize_batch(eng._h, b.data.data_ptr(), b.offsets.data_ptr(), 2, 0, b.end, 1, i32.data_ptr(),
                                          i32.data_ptr(), 1, sp.data_ptr(), None, res.data_ptr(), ws.data_ptr(), ws.numel(), None) == C.OK
    r = res.cpu()
    assert int(r[0]) == 2 and int(r[2]) & C.ST_OVERFLOW
    # the Python layer maps them onto the reference's exceptions (tokenizer.py:88-102)
    with pytest.raises(RuntimeError):
        A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'))            # a BPE JSON loaded as sentencepiece
    with pytest.raises(Exception):
        A.aksharTokenizer(os.path.join(models_dir, 'spm24k.model'), 'bpe')
    assert A.aksharTokenizer('/nonexistent/model').model_type == 'akshar'      # silent fallback, like the reference


def test_sharded_encode_two_gpus(A):
    """the N > 1 entry point on real GPUs (skipped on a one-GPU box): two ranks under torchrun, gathered ids == one GPU's"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                          '--master-port', '29533', os.path.join(root, 'tools', 'sharded_check.py')], capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count(': True') == 2


def test_invalid_utf8_is_survived(A, eng, models_dir):
    """the reference only ever sees Python strings; the C ABI takes bytes.  Arbitrary bytes (stray continuation bytes,
    truncated sequences, 0xFF, overlong forms) must come back as a status or as garbage -- never as a fault or a hang"""
    import torch
    from akshar_b200 import _lib as C
    rng = np.random.default_rng(99)
    tb = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe')
    tu = A.aksharTokenizer(os.path.join(models_dir, 'spm24k.model'), 'sentencepiece')
    pools = [np.arange(256, dtype=np.uint8),
             np.array([0x80, 0xBF, 0xC0, 0xC2, 0xE0, 0xA4, 0xA5, 0xBC, 0x8D, 0xED, 0xF0, 0xF4, 0xFF, 0x20, 0x61, 0x0A, 0x3C, 0x73, 0x3E], dtype=np.uint8)]
    for pool in pools:
        for n_rows, max_len in ((3000, 200), (20, 20000), (1, 300000)):
            lens = rng.integers(0, max_len + 1, size=n_rows)
            off = np.zeros(n_rows + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            data = pool[rng.integers(0, pool.size, size=int(off[-1]))]
            host = (torch.from_numpy(data), torch.from_numpy(off))
            for e in (eng, tb._eng, tu._eng):
                b = e.put(host)
                e.normalize_batch(b, check=False)
                e.normalize_batch(b, clean_hinglish=False, check=False)
                e.segment_batch(b, clusters=True, runs=True, check=False)
                e.segment_masks(b, clusters=True, runs=True, check=False)
                e.normalize_segment_batch(b, check=False)
                for rule in (C.WORDS_HINDI, C.WORDS_SPLIT):
                    try:
                        e.word_tokenize_batch(b, rule=rule, row_flags=True)
                    except Exception as ex:
                        assert 'word_tokenize' in str(ex)
                try:
                    e.lines_batch(b.data, b.n_bytes)
                except Exception as ex:
                    assert 'lines' in str(ex)
            for tk, kind in ((tb, 0), (tu, 1)):
                for mode in (C.MODE_TILES, C.MODE_ROWS):
                    tk._eng.tokenizer_encode_batch(host, kind, mode=mode, check=False)
                    tk._eng.tokenizer_encode_batch(host, kind, clean_hinglish=False, mode=mode, check=False)
                # ids that are no ids: negative, beyond the vocabulary, as int32
                ids = torch.from_numpy(rng.integers(-5, 40000, size=int(off[-1]) // 4 + 1).astype(np.int32))
                sp = torch.from_numpy(np.minimum(off // 4, ids.numel()).astype(np.int64))
                for form in (C.FORM_DECODE, C.FORM_DETOKENIZE):
                    try:
                        tk._eng.decode_batch(ids, sp, kind, form)
                    except IndexError:
                        assert kind == 1                              # SentencePiece: 'piece id is out of range.'
            try:
                eng.signature_batch(eng.put(host))
            except Exception as ex:
                assert 'signature' in str(ex)
            torch.cuda.synchronize()
    # the contexts are still good for real work afterwards
    assert tb.encode('hello') == O.bpe_encode(O.BpeModel(os.path.join(models_dir, 'bpe24k.json')), 'hello')
    assert tu.decode(tu.encode('\u0928\u092e\u0938\u094d\u0924\u0947 world')) == '\u0928\u092e\u0938\u094d\u0924\u0947 world'


def test_texts_that_do_not_start_at_an_aligned_zero(A, eng, golden):
    """text_begin != 0 and addresses that are not 16-byte aligned through the round-2 entry points: same results"""
    import torch
    from akshar_b200 import _lib as C
    lines = [r['in'] for r in golden['rows'][:1500]] + ['', 'x']
    data, off = sc.pack(lines)
    base = eng.put((torch.from_numpy(data), torch.from_numpy(off)))
    nb = eng.normalize_batch(base)
    wb0, we0, sp0, fl0 = eng.word_tokenize_batch(nb, rule=C.WORDS_HINDI, row_flags=True)
    n0, mk0 = eng.normalize_segment_batch(base)
    for lead in (1, 5, 16, 21):
        pad = np.full(lead, 0x61, dtype=np.uint8)
        d = torch.from_numpy(np.concatenate([pad, data])).cuda()
        o = torch.from_numpy(off + lead).cuda()
        tb = A.TextBatch(d, o, lead, lead + int(off[-1]))
        nd = torch.cat([torch.from_numpy(pad).cuda(), nb.data[:nb.end]])
        tn = A.TextBatch(nd, nb.offsets + lead, lead, lead + nb.end)
        wb, we, sp, fl = eng.word_tokenize_batch(tn, rule=C.WORDS_HINDI, row_flags=True)
        assert torch.equal(wb, wb0) and torch.equal(we, we0) and torch.equal(sp, sp0) and torch.equal(fl, fl0)
        n1, mk1 = eng.normalize_segment_batch(tb)
        assert n1.end == n0.end and torch.equal(n1.data[:n1.end], n0.data[:n0.end]) and torch.equal(n1.offsets, n0.offsets)
        W = (n0.end + 32) // 32
        for key in ('cluster', 'run'):
            assert torch.equal(mk1[key][:W], mk0[key][:W])
        assert torch.equal(mk1['tags'][:, :W], mk0['tags'][:, :W])
        # a file whose first byte sits at an odd address
        f = eng.join_rows(base)
        shifted = torch.cat([torch.from_numpy(pad).cuda(), f])[lead:]
        rows_a, rows_b = eng.lines_batch(f, f.numel()), eng.lines_batch(shifted, shifted.numel())
        assert rows_a.end == rows_b.end and torch.equal(rows_a.offsets, rows_b.offsets)
        assert torch.equal(rows_a.data[:rows_a.end], rows_b.data[:rows_b.end])


def test_edge_batches_round2_entries(A, eng, models_dir, tmp_path):
    """no rows, empty rows, one-byte rows and files of nothing but line ends through the word tokenizers, decode,
    composition, cluster merges, the fused normalize + segment call and the file front end"""
    from akshar_b200 import corpus, segment as S
    for fn in (S.word_tokenize_hindi_batch, S.word_tokenize_batch, S.akshara_level_tokenization_batch, S.preserve_nukta_batch,
               S.analyze_text_composition_batch):
        assert fn([]) == []
    assert S.normalize_and_segment_batch([]) == ([], [])
    edge = ['', ' ', '\u0964', 'a', '', '\u093c', '\u0915\u094d', '.', '\u3000', '']
    assert S.word_tokenize_hindi_batch(edge) == [O.word_tokenize_hindi(t) for t in edge]
    assert S.word_tokenize_batch(edge) == [O.word_tokenize(t) for t in edge]
    assert S.akshara_level_tokenization_batch(edge) == [O.akshara_level_tokenization(t) for t in edge]
    assert S.preserve_nukta_batch(edge) == [O.preserve_nukta(t) for t in edge]
    assert S.analyze_text_composition_batch(edge) == [O.analyze_text_composition(t) for t in edge]
    strings, akshars = S.normalize_and_segment_batch(edge)
    assert strings == [O.normalize_text(t) for t in edge]
    assert akshars == [O.segment_akshars(O.normalize_text(t)) for t in edge]
    for name, kind in (('bpe24k.json', 'bpe'), ('spm24k.model', 'sentencepiece')):
        tk = A.aksharTokenizer(os.path.join(models_dir, name), kind)
        assert tk.decode_batch([]) == [] and tk.detokenize_batch([]) == []
        assert tk.decode_batch([[], [], []]) == ['', '', ''] and tk.detokenize_batch([[], []]) == ['', '']
        one = tk.encode('a')
        assert tk.decode_batch([[], one, []]) == ['', tk.decode(one), '']
    for i, body in enumerate((b'', b'\n', b'\r\n\r\n', b' \t \n\n  ', b'x', b'\nx', b'x\n', b'\r', b' x \r y ')):
        f = tmp_path / f'e{i}.txt'
        f.write_bytes(body)
        with open(f, 'r', encoding='utf-8') as h:
            ref = [ln.strip() for ln in h.readlines() if ln.strip()]
        assert corpus.read_rows(str(f)) == ref, body


def test_normalize_text_that_expands(A, eng):
    """NFC makes composition exclusions longer (U+0958-095F: 3 -> 6 bytes, U+FB2A-FB4E: 3 -> 4-6 bytes): warp tiles full
    of them overflow the writer's shared-memory stage (the direct path) and nearly every lane is a slow lane.  Runs of
    such letters are kept below the 64 code points one NFC segment may have (AKSHAR_ST_NFC_SEGMENT, checked last)."""
    import random
    from akshar_b200 import segment as S
    from akshar_b200.batch import BatchStatusError
    rng = random.Random(5)
    excl = [chr(c) for c in range(0x958, 0x960)] + ['\ufb2a', '\ufb2b', '\ufb2e', '\ufb4b', '\u0f43', '\u2adc']
    lines = []
    for i in range(600):
        n = rng.choice((1, 7, 40, 200, 700, 3000))
        body = []
        for k in range(n):
            body.append(rng.choice('\u0915\u093e\u0964') if k % 20 == 19 else rng.choice(excl) if rng.random() < 0.95 else rng.choice('ab '))
        lines.append(''.join(body))
    lines += ['\u0958\u093e' * 2500, 'x' + '\u095c\u094d\u0915' * 1700, '']
    for nr, nc in ((True, True), (True, False), (False, False)):
        got = A.normalize_batch(lines, normalize_roman=nr, clean_hinglish=nc)
        assert got == [O.normalize_text(t, nr, nc) for t in lines]
    strings, akshars = S.normalize_and_segment_batch(lines)
    assert strings == [O.normalize_text(t) for t in lines]
    assert akshars == [O.segment_akshars(t) for t in strings]
    with pytest.raises(BatchStatusError):       # 127+ such letters in a row are one NFC segment of more than 256 code points: refused loudly
        A.normalize_batch(['\u0958' * 200])


def test_zalgo_text(A, eng, models_dir):
    """stacked combining marks (social text has them): up to ~250 marks on one base are reordered / composed exactly as
    the reference does; the encoders see the same rows"""
    import random
    rng = random.Random(11)
    marks = [chr(c) for c in (0x300, 0x301, 0x302, 0x303, 0x308, 0x30a, 0x316, 0x317, 0x323, 0x324, 0x325, 0x327, 0x328, 0x32d,
                              0x334, 0x335, 0x336, 0x338, 0x340, 0x341, 0x343, 0x344, 0x345, 0x35c, 0x360, 0x489, 0x93c, 0x94d,
                              0x951, 0x952)]
    lines = []
    for i in range(400):
        parts = []
        for _ in range(rng.choice((1, 3, 8))):
            base = rng.choice('aeou AEH\u0915\u0930\u0928z')
            k = rng.choice((0, 1, 5, 20, 45, 70, 120, 200, 240))
            parts.append(base + ''.join(rng.choice(marks) for _ in range(k)))
        lines.append(' '.join(parts))
    for nr, nc in ((True, True), (True, False), (False, False)):
        assert A.normalize_batch(lines, normalize_roman=nr, clean_hinglish=nc) == [O.normalize_text(t, nr, nc) for t in lines]
    norm = [O.normalize_text(t) for t in lines]
    assert A.segment_akshars_batch(norm) == [O.segment_akshars(t) for t in norm]
    tk = A.aksharTokenizer(os.path.join(models_dir, 'bpe24k.json'), 'bpe')
    om = O.BpeModel(os.path.join(models_dir, 'bpe24k.json'))
    assert tk.encode_batch(lines) == [O.bpe_encode(om, t) for t in norm]
    tu = A.aksharTokenizer(os.path.join(models_dir, 'spm24k.model'))
    ou = O.UnigramModel(os.path.join(models_dir, 'spm24k.model'))
    assert tu.encode_batch(lines[:150]) == [O.unigram_encode(ou, t) for t in norm[:150]]
