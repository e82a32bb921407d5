"""CPU oracle for the Akshar hot path -- TEST INFRASTRUCTURE ONLY.

This file is a from-the-spec restatement of what the reference computes on the path
normalize -> grapheme clusters -> script runs -> BPE / Unigram encode.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` leg may import it;
the product package `akshar_b200` never does (it fails loudly without its CUDA library).

The reference (`/root/reference/src/akshar`) is pure Python and delegates the arithmetic to third-party
engines that are NOT vendored in the reference tree.  Each function below restates the PUBLISHED algorithm
of that engine, with property tables probed from the pinned versions by tools/gen_tables.py:

  regex 2026.3.32 (Unicode 17.0)      `\\X`  -> UAX #29 extended grapheme clusters (GB1-GB999, GB9c, GB11)
  CPython 3.12.3 unicodedata (15.0)   NFC   -> UAX #15 (decompose, canonical reorder, compose)
  tokenizers 0.22.2                   BPE   -> NFKC, `\\w+|[^\\w\\s]+`, lowest-rank-leftmost merges, <s> $A </s>
  sentencepiece 0.2.1                 Unigram -> whitespace escape, f32 Viterbi, strict `>`, byte fallback

Parity pin: tests/test_oracle_golden.py checks every function here against tests/golden/*.json, which were
produced by running the UNMODIFIED reference (imported from /root/reference/src in the dev container) through
tools/make_golden.py; the reference's own test expectations (tests/test_normalize.py, test_segment.py) and the
executed notebook outputs are included in those vectors.
"""
import json
import os
import struct

_HERE = os.path.dirname(os.path.abspath(__file__))
NCP = 0x110000

GCB_OTHER, GCB_CR, GCB_LF, GCB_CONTROL, GCB_EXTEND, GCB_ZWJ, GCB_RI, GCB_PREPEND, GCB_SPACINGMARK, \
    GCB_L, GCB_V, GCB_T, GCB_LV, GCB_LVT = range(14)
INCB_NONE, INCB_CONSONANT, INCB_LINKER, INCB_EXTEND = range(4)
TAGS = ['devanagari', 'roman', 'digit', 'punct', 'other']


class _Tables:
    def __init__(self):
        with open(os.path.join(_HERE, 'ucd_tables.json')) as f:
            j = json.load(f)
        self.versions = j['versions']

        def vals(name):
            a = bytearray(NCP)
            for lo, hi, v in j[name]:
                a[lo:hi + 1] = bytes([v]) * (hi - lo + 1)
            return a

        def flags(name):
            a = bytearray(NCP)
            for lo, hi in j[name]:
                a[lo:hi + 1] = b'\x01' * (hi - lo + 1)
            return a

        self.gcb = vals('gcb')
        self.incb = vals('incb')
        self.extpict = flags('extpict')
        self.allow = flags('allow')
        self.isdigit = flags('isdigit')
        self.ccc = vals('ccc')
        self.nfc_qc = vals('nfc_qc')
        self.hf_class = vals('hf_class')
        self.bpe_safe = flags('bpe_safe')
        self.decomp = {int(k, 16): v for k, v in j['decomp'].items()}
        self.pairs = {(a, b): c for a, b, c in j['pairs']}
        self.latin_lower = {int(k, 16): v for k, v in j['latin_lower'].items()}
        self.full_lower = {int(k, 16): v for k, v in j['full_lower'].items()}
        self.hf_kmap = {int(k, 16): v for k, v in j['hf_kmap'].items()}
        self.hf_unknown = frozenset(j['hf_unknown'])


_T = None


def tables():
    global _T
    if _T is None:
        _T = _Tables()
    return _T


# --------------------------------------------------------------------------------------------------
# a1  normalize_unicode  (reference normalize.py:13-18 -> unicodedata.normalize('NFC'), UAX #15)
# --------------------------------------------------------------------------------------------------
_SB, _LB, _VB, _TB = 0xAC00, 0x1100, 0x1161, 0x11A7
_LC, _VC, _TC = 19, 21, 28
_NC, _SC = _VC * _TC, _LC * _VC * _TC


def _decompose(cp, out):
    T = tables()
    if _SB <= cp < _SB + _SC:
        s = cp - _SB
        out.append(_LB + s // _NC)
        out.append(_VB + (s % _NC) // _TC)
        t = s % _TC
        if t:
            out.append(_TB + t)
        return
    d = T.decomp.get(cp)
    if d is None:
        out.append(cp)
    else:
        out.extend(d)


def _compose_pair(a, b):
    if _LB <= a < _LB + _LC and _VB <= b < _VB + _VC:
        return _SB + ((a - _LB) * _VC + (b - _VB)) * _TC
    if _SB <= a < _SB + _SC and (a - _SB) % _TC == 0 and _TB < b < _TB + _TC:
        return a + (b - _TB)
    return tables().pairs.get((a, b))


def nfc_cps(cps):
    """UAX #15 NFC over a list of code points."""
    T = tables()
    ccc = T.ccc
    d = []
    for cp in cps:
        _decompose(cp, d)
    # canonical ordering: stable sort of every maximal run of non-starters by ccc
    i, n = 0, len(d)
    while i < n:
        if ccc[d[i]] == 0:
            i += 1
            continue
        j = i
        while j < n and ccc[d[j]] != 0:
            j += 1
        d[i:j] = sorted(d[i:j], key=lambda c: ccc[c])
        i = j
    # canonical composition
    out = []
    starter = -1      # index in out of the last starter
    last_ccc = -1     # ccc of the last char appended after the starter (-1: none)
    for c in d:
        cc = ccc[c]
        # c is unblocked iff it is adjacent to the starter or the last kept mark has a strictly lower ccc
        if starter >= 0 and (last_ccc == -1 or last_ccc < cc):
            comp = _compose_pair(out[starter], c)
            if comp is not None:
                out[starter] = comp
                continue
        if cc == 0:
            starter = len(out)
            last_ccc = -1
        else:
            last_ccc = cc
        out.append(c)
    return out


def normalize_unicode(text):
    return ''.join(map(chr, nfc_cps([ord(c) for c in text])))


# --------------------------------------------------------------------------------------------------
# a2-a5  semantic_normalize / filter_garbage / remove_elongations / normalize_text
#        (reference normalize.py:21-56, 92-148; algorithm statement SURVEY.md B1)
# --------------------------------------------------------------------------------------------------
def semantic_normalize_cps(cps):
    ll = tables().latin_lower
    out = []
    for c in cps:
        m = ll.get(c)
        if m is None:
            out.append(c)
        else:
            out.extend(m)
    return out


def filter_garbage_cps(cps):
    allow = tables().allow
    return [c for c in cps if allow[c]]


def remove_elongations_cps(cps):
    """runs of >= 3 identical code points collapse to one; `.` does not match U+000A"""
    out = []
    i, n = 0, len(cps)
    while i < n:
        j = i
        while j < n and cps[j] == cps[i]:
            j += 1
        run = j - i
        if run >= 3 and cps[i] != 0x0A:
            out.append(cps[i])
        else:
            out.extend(cps[i:j])
        i = j
    return out


def semantic_normalize(text):
    return ''.join(map(chr, semantic_normalize_cps([ord(c) for c in text])))


def filter_garbage(text):
    return ''.join(map(chr, filter_garbage_cps([ord(c) for c in text])))


def remove_elongations(text):
    return ''.join(map(chr, remove_elongations_cps([ord(c) for c in text])))


def normalize_hinglish(text):
    return remove_elongations(filter_garbage(text))


def normalize_text(text, normalize_roman=True, clean_hinglish=True):
    cps = nfc_cps([ord(c) for c in text])
    if normalize_roman:
        cps = semantic_normalize_cps(cps)
    if clean_hinglish:
        cps = remove_elongations_cps(filter_garbage_cps(cps))
    return ''.join(map(chr, cps))


# --------------------------------------------------------------------------------------------------
# a6  roman_phonetic_signature  (reference normalize.py:59-89)
# --------------------------------------------------------------------------------------------------
def _lower_full(text):
    # str.lower(): simple mapping, U+0130 -> i + U+0307, and the Final_Sigma context rule.
    # cased / case-ignorable are not in the JSON; Final_Sigma is resolved with str methods on single chars only.
    fl = tables().full_lower
    out = []
    n = len(text)
    for i, ch in enumerate(text):
        c = ord(ch)
        if c == 0x3A3:
            out.append(_final_sigma(text, i))
            continue
        m = fl.get(c)
        if m is None:
            out.append(ch)
        else:
            out.extend(chr(x) for x in m)
    return ''.join(out)


def _case_ignorable(ch):
    # probe of CPython's _PyUnicode_IsCaseIgnorable through a 3-char lower() (single-char context only)
    return ('aΣ' + ch).lower()[1] == 'ς' and ('aΣ' + ch + 'a').lower()[1] == 'σ'


def _cased(ch):
    return ('aΣ' + ch).lower()[1] == 'σ'


def _final_sigma(text, i):
    j = i - 1
    while j >= 0 and _case_ignorable(text[j]):
        j -= 1
    if j < 0 or not _cased(text[j]):
        return 'σ'
    j = i + 1
    while j < len(text) and _case_ignorable(text[j]):
        j += 1
    if j == len(text) or not _cased(text[j]):
        return 'ς'
    return 'σ'


def roman_phonetic_signature(word):
    w = _lower_full(word)
    w = remove_elongations(w)
    # `$` matches at the very end or before a final '\n'
    for pat, rep in (('ee', 'i'), ('oo', 'u')):
        if w.endswith(pat):
            w = w[:-2] + rep
        elif w.endswith(pat + '\n'):
            w = w[:-3] + rep + '\n'
    for pat, rep in (('aa', 'a'), ('kh', 'k'), ('gh', 'g'), ('ch', 'c'), ('th', 't'), ('ph', 'p'), ('bh', 'b'),
                     ('dh', 'd')):
        w = w.replace(pat, rep)
    return w


# --------------------------------------------------------------------------------------------------
# a7  segment_akshars -> regex `\X`  (reference segment.py:14,78; UAX #29 rules)
# --------------------------------------------------------------------------------------------------
def grapheme_breaks(cps):
    """returns list of cluster END indices (in code points), i.e. positions i where a boundary lies before cps[i],
    plus len(cps).  Empty input -> []."""
    T = tables()
    gcb, incb, ext = T.gcb, T.incb, T.extpict
    n = len(cps)
    if n == 0:
        return []
    ends = []
    # running context
    ri_run = 0              # number of consecutive RI immediately before position i
    conj = 0                # GB9c state: 0 none, 1 seen Consonant [Extend|Linker]* without Linker, 2 .. with Linker
    pict = 0                # GB11 state: 0 none, 1 ExtPict Extend*, 2 ExtPict Extend* ZWJ
    for i in range(n):
        b = cps[i]
        gb = gcb[b]
        if i > 0:
            a = cps[i - 1]
            ga = gcb[a]
            if ga == GCB_CR and gb == GCB_LF:
                brk = False                                         # GB3
            elif ga in (GCB_CONTROL, GCB_CR, GCB_LF) or gb in (GCB_CONTROL, GCB_CR, GCB_LF):
                brk = True                                          # GB4, GB5
            elif ga == GCB_L and gb in (GCB_L, GCB_V, GCB_LV, GCB_LVT):
                brk = False                                         # GB6
            elif ga in (GCB_LV, GCB_V) and gb in (GCB_V, GCB_T):
                brk = False                                         # GB7
            elif ga in (GCB_LVT, GCB_T) and gb == GCB_T:
                brk = False                                         # GB8
            elif gb in (GCB_EXTEND, GCB_ZWJ):
                brk = False                                         # GB9
            elif gb == GCB_SPACINGMARK:
                brk = False                                         # GB9a
            elif ga == GCB_PREPEND:
                brk = False                                         # GB9b
            elif conj == 2 and incb[b] == INCB_CONSONANT:
                brk = False                                         # GB9c
            elif pict == 2 and ext[b]:
                brk = False                                         # GB11
            elif ga == GCB_RI and gb == GCB_RI and (ri_run & 1) == 1:
                brk = False                                         # GB12, GB13
            else:
                brk = True                                          # GB999
            if brk:
                ends.append(i)
        # advance context over b
        ri_run = ri_run + 1 if gb == GCB_RI else 0
        ib = incb[b]
        if ib == INCB_CONSONANT:
            conj = 1
        elif ib == INCB_LINKER:
            conj = 2 if conj else 0
        elif ib == INCB_EXTEND:
            pass
        else:
            conj = 0
        if ext[b]:
            pict = 1
        elif gb == GCB_EXTEND:
            pict = 1 if pict == 1 else 0
        elif gb == GCB_ZWJ:
            pict = 2 if pict == 1 else 0
        else:
            pict = 0
    ends.append(n)
    return ends


_MATRA_RANGES = ((0x0900, 0x0902), (0x093E, 0x094C), (0x0951, 0x0954))


def _is_matra_cp(cp):
    return any(lo <= cp <= hi for lo, hi in _MATRA_RANGES)


def segment_breaks(cps, matras=False):
    """cluster (or cluster-part, for matras=True) END indices in code points (reference segment.py:40-125)"""
    ends = grapheme_breaks(cps)
    if not matras:
        return ends
    out = []
    start = 0
    for e in ends:
        # inside one cluster every matra / halant is its own part; other cps accumulate
        have_base = False
        for i in range(start, e):
            c = cps[i]
            if _is_matra_cp(c) or c == 0x094D:
                if have_base:
                    out.append(i)
                    have_base = False
                out.append(i + 1)
            else:
                have_base = True
        if have_base:
            out.append(e)
        start = e
    return out


def segment_akshars(text, matras=False, separate_matras=None):
    if separate_matras is not None:
        matras = separate_matras
    cps = [ord(c) for c in text]
    ends = segment_breaks(cps, matras)
    out = []
    s = 0
    for e in ends:
        out.append(text[s:e])
        s = e
    return out


# --------------------------------------------------------------------------------------------------
# a9-a11  identify_script / detect_code_switches / analyze_text_composition (reference segment.py:128-236; B3)
# --------------------------------------------------------------------------------------------------
_PUNCT = set(ord(c) for c in ' .,!?;:\'"()-[]{}')


def script_tag(cp):
    if 0x0900 <= cp <= 0x097F:
        return 0
    if 0x41 <= cp <= 0x5A or 0x61 <= cp <= 0x7A:
        return 1
    if tables().isdigit[cp]:
        return 2
    if cp in _PUNCT:
        return 3
    return 4


def identify_script(char):
    return TAGS[script_tag(ord(char))]


def script_runs(cps):
    """-> list of (end_index_in_cps, tag or None)"""
    runs = []
    cur = None
    for i, c in enumerate(cps):
        t = script_tag(c)
        if t in (2, 3):
            continue
        if cur is None:
            cur = t
        elif t != cur:
            runs.append((i, cur))
            cur = t
    if cps:
        runs.append((len(cps), cur))
    return runs


def detect_code_switches(text):
    cps = [ord(c) for c in text]
    out = []
    s = 0
    for e, t in script_runs(cps):
        out.append((text[s:e], None if t is None else TAGS[t]))
        s = e
    return out


def segment_by_script(text):
    return [s for s, _ in detect_code_switches(text)]


def analyze_text_composition(text):
    cps = [ord(c) for c in text]
    ak = grapheme_breaks(cps)
    runs = script_runs(cps)
    total = len(cps)
    dev = rom = 0
    s = 0
    for e, t in runs:
        if t == 0:
            dev += e - s
        elif t == 1:
            rom += e - s
        s = e
    return {
        'akshar_count': len(ak),
        'script_switches': len(runs) - 1,
        'devanagari_ratio': dev / total if total > 0 else 0,
        'roman_ratio': rom / total if total > 0 else 0,
    }


# --------------------------------------------------------------------------------------------------
# f1  the rows of a text file (reference cli.py:165-190 = scripts/train_bpe.py:16-35 = scripts/train_spm.py:18-44)
# --------------------------------------------------------------------------------------------------
def file_rows(data):
    """`open(path, 'r', encoding='utf-8').readlines()`, `strip()`, empty lines skipped -- from the file's bytes.  A text-mode
    file translates '\\r\\n' and a lone '\\r' to '\\n' (universal newlines) and readlines() cuts after every '\\n'."""
    text = bytes(data).decode('utf-8').replace('\r\n', '\n').replace('\r', '\n')
    rows = []
    for line in text.split('\n'):
        k = len(line)
        i = 0
        while i < k and ord(line[i]) in _PY_SPACE:
            i += 1
        while k > i and ord(line[k - 1]) in _PY_SPACE:
            k -= 1
        if k > i:
            rows.append(line[i:k])
    return rows


# --------------------------------------------------------------------------------------------------
# f4  the feature wrappers that are functions of the cluster boundaries (reference features.py:28-55, 173-206)
# --------------------------------------------------------------------------------------------------
def akshara_level_tokenization(text):
    """features.py:28-55: clusters that hold a halant pile up, a cluster without one flushes the pile and stands alone"""
    out = []
    pile = []
    for c in segment_akshars(text):
        if '\u094d' in c:
            pile.append(c)
        else:
            if pile:
                out.append(''.join(pile))
                pile = []
            out.append(c)
    if pile:
        out.append(''.join(pile))
    return out


def preserve_nukta(text):
    """features.py:173-206: a cluster that holds U+093C is joined with the cluster after it (which is then skipped)"""
    seg = segment_akshars(text)
    out = []
    i = 0
    while i < len(seg):
        if '\u093c' in seg[i] and i + 1 < len(seg):
            out.append(seg[i] + seg[i + 1])
            i += 2
        else:
            out.append(seg[i])
            i += 1
    return out


# --------------------------------------------------------------------------------------------------
# f3  word tokenizers (reference segment.py:239-401)
# --------------------------------------------------------------------------------------------------
# str.isspace() (CPython, Unicode 15.0: bidirectional class WS / B / S or category Zs), written out
_PY_SPACE = frozenset([0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x1C, 0x1D, 0x1E, 0x1F, 0x20, 0x85, 0xA0, 0x1680] + list(range(0x2000, 0x200B)) +
                      [0x2028, 0x2029, 0x202F, 0x205F, 0x3000])
_WORD_PUNCT = frozenset(ord(c) for c in '.,!?;:()[]{}"\'')      # segment.py:272 `other_punct`


def _word_loop(cps):
    """segment.py:270-297 (Hindi) == segment.py:335-362 (Sanskrit): -> list of (begin, end) code point ranges"""
    out = []
    start = -1
    for i, cp in enumerate(cps):
        if cp in _PY_SPACE or cp in _WORD_PUNCT:
            if start >= 0:
                out.append((start, i))
                start = -1
        elif cp == 0x0964 or cp == 0x0965:
            if start >= 0:
                out.append((start, i))
                start = -1
            out.append((i, i + 1))
        elif start < 0:
            start = i
    if start >= 0:
        out.append((start, len(cps)))
    return out


def _split_loop(cps):
    """str.split(): runs of non-space code points"""
    out = []
    start = -1
    for i, cp in enumerate(cps):
        if cp in _PY_SPACE:
            if start >= 0:
                out.append((start, i))
                start = -1
        elif start < 0:
            start = i
    if start >= 0:
        out.append((start, len(cps)))
    return out


def word_tokenize_hindi(text, use_morphology=False):
    """segment.py:239-300 without a Morfessor model (none ships with the reference: morph.py falls through to this loop)"""
    n = normalize_text(text, normalize_roman=True, clean_hinglish=True)
    return [n[b:e] for b, e in _word_loop([ord(c) for c in n])]


def word_tokenize_sanskrit(text, use_morphology=False):
    """segment.py:303-362"""
    return word_tokenize_hindi(text, use_morphology)


def word_tokenize(text, language='auto', use_morphology=False):
    """segment.py:365-401"""
    if language == 'auto':
        if any(0x0900 <= ord(c) <= 0x097F for c in text):
            language = 'hindi'
        else:
            return [text[b:e] for b, e in _split_loop([ord(c) for c in text])]
    if language.lower() in ('hindi', 'hi', 'hin'):
        return word_tokenize_hindi(text, use_morphology)
    if language.lower() in ('sanskrit', 'sa', 'san', 'skr'):
        return word_tokenize_sanskrit(text, use_morphology)
    return [text[b:e] for b, e in _split_loop([ord(c) for c in text])]


# --------------------------------------------------------------------------------------------------
# a17  BPE (HuggingFace tokenizers JSON written by the reference's scripts/train_bpe.py:68-98)
# --------------------------------------------------------------------------------------------------
class BpeModel:
    """Parsed from the tokenizer JSON with the json module only (no `tokenizers` import)."""

    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray)):
            j = json.loads(bytes(path_or_bytes).decode('utf-8'))
        else:
            with open(path_or_bytes, 'rb') as f:
                j = json.loads(f.read().decode('utf-8'))
        m = j['model']
        assert m['type'] == 'BPE'
        self.vocab = dict(m['vocab'])
        self.id_to_token = {}
        for t, i in self.vocab.items():
            self.id_to_token[i] = t
        self.merges = {}
        for rank, mg in enumerate(m['merges']):
            a, b = mg.split(' ') if isinstance(mg, str) else mg
            self.merges[(self.vocab[a], self.vocab[b])] = (rank, self.vocab[a + b])
        self.specials = [(t['content'], t['id']) for t in j.get('added_tokens', [])]
        self.special_ids = set(i for _, i in self.specials)
        self.special_first = frozenset(c[0] for c, _ in self.specials if c)
        for c, i in self.specials:
            self.id_to_token[i] = c
        self.bos = self.eos = None
        pp = j.get('post_processor')
        if pp and pp.get('type') == 'TemplateProcessing':
            st = pp['special_tokens']
            single = pp['single']
            first = single[0].get('SpecialToken')
            last = single[-1].get('SpecialToken')
            if first:
                self.bos = st[first['id']]['ids'][0]
            if last:
                self.eos = st[last['id']]['ids'][0]
        self.normalizer = (j.get('normalizer') or {}).get('type')
        self.pre_tokenizer = (j.get('pre_tokenizer') or {}).get('type')
        self.unk = m.get('unk_token')

    def vocab_size(self):
        return len(set(self.vocab.values()) | self.special_ids)


_NFKC_SPACE = {0xA0, 0x2002, 0x2003, 0x2004, 0x2005, 0x2006, 0x2007, 0x2008, 0x2009, 0x200A, 0x202F, 0x205F, 0x3000}


def bpe_word_ids(model, word_cps):
    """HF `Word::merge_all`: repeatedly merge the adjacent pair of lowest rank (leftmost on ties)."""
    sym = []
    for c in word_cps:
        i = model.vocab.get(chr(c))
        if i is not None:          # unk_token is null -> characters outside the vocab are dropped
            sym.append(i)
    while len(sym) > 1:
        best = None
        for i in range(len(sym) - 1):
            r = model.merges.get((sym[i], sym[i + 1]))
            if r is not None and (best is None or r[0] < best[0]):
                best = (r[0], i, r[1])
        if best is None:
            break
        _, i, nid = best
        sym[i:i + 2] = [nid]
    return sym


def hf_nfkc_cps(cps):
    """`normalizers.NFKC()` of tokenizers 0.22.2 (scripts/train_bpe.py:71) over a list of code points.  Its Unicode data is
    older than CPython's: code points it does not know (tables().hf_unknown) pass through as inert starters, every other
    code point is replaced by HF's compatibility decomposition (tables().hf_kmap, probed) or the canonical one, then
    canonical ordering and composition as in NFC; a composite HF does not know is left decomposed."""
    T = tables()
    out = []
    run = []

    def flush():
        if run:
            for c in nfc_cps(run):
                if c in T.hf_unknown:
                    _decompose(c, out)
                else:
                    out.append(c)
            del run[:]
    for c in cps:
        if c in T.hf_unknown:
            flush()
            out.append(c)
        else:
            run.extend(T.hf_kmap.get(c, (c,)))
    flush()
    return out


def _split_specials(model, text):
    """HF AddedVocabulary: the special tokens (normalized=false) are cut out of the RAW text, leftmost-longest, before the
    normalizer runs (scripts/train_bpe.py:80) -> [(text piece, None) | (None, special id)]"""
    specials = sorted(model.specials, key=lambda x: -len(x[0]))
    out = []
    i = start = 0
    n = len(text)
    while i < n:
        hit = None
        if text[i] in model.special_first:
            for c, sid in specials:
                if text.startswith(c, i):
                    hit = (c, sid)
                    break
        if hit is None:
            i += 1
            continue
        if i > start:
            out.append((text[start:i], None))
        out.append((None, hit[1]))
        i += len(hit[0])
        start = i
    if start < n:
        out.append((text[start:], None))
    return out


def bpe_encode(model, text):
    """ids of `Tokenizer.encode(text).ids` (tokenizer.py:193): special tokens cut out of the raw text, every other piece
    through NFKC -> Whitespace pre-tokenizer (`\\w+|[^\\w\\s]+`) -> BPE merges, framed by the template's <s> ... </s>"""
    T = tables()
    hf = T.hf_class
    ids = []
    for piece, sid in _split_specials(model, text):
        if sid is not None:
            ids.append(sid)
            continue
        cps = hf_nfkc_cps([ord(c) for c in piece])
        i, n = 0, len(cps)
        while i < n:
            k = hf[cps[i]]
            if k == 2:
                i += 1
                continue
            j = i
            while j < n and hf[cps[j]] == k:
                j += 1
            ids.extend(bpe_word_ids(model, cps[i:j]))
            i = j
    if model.bos is not None:
        ids = [model.bos] + ids
    if model.eos is not None:
        ids = ids + [model.eos]
    return ids


def bpe_decode(model, ids):
    """tokenizer.py:219-220 -> HF Tokenizer.decode(ids) with `decoder: null`: the tokens of the ids that are in the vocabulary
    and not special, joined by one space (tokenizers 0.22.2; ids without a token are skipped -- tests/golden decode_fuzz)"""
    return ' '.join(model.id_to_token[i] for i in ids if i in model.id_to_token and i not in model.special_ids)


def bpe_detokenize(tokens):
    """tokenizer.py:240-244"""
    return ' '.join(tokens).replace(' ##', '').replace('\u0120', ' ').strip()


def spm_detokenize(tokens):
    """tokenizer.py:236-239"""
    return ''.join(tokens).replace('\u2581', ' ').strip()


# --------------------------------------------------------------------------------------------------
# a18  Unigram (SentencePiece ModelProto written by scripts/train_spm.py:80-108)
# --------------------------------------------------------------------------------------------------
def _pb_fields(buf):
    """minimal protobuf wire-format reader: yields (field_no, wire_type, value)"""
    i, n = 0, len(buf)
    while i < n:
        key = 0
        shift = 0
        while True:
            b = buf[i]
            i += 1
            key |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v = 0
            shift = 0
            while True:
                b = buf[i]
                i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
        elif wt == 1:
            v = buf[i:i + 8]
            i += 8
        elif wt == 2:
            ln = 0
            shift = 0
            while True:
                b = buf[i]
                i += 1
                ln |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            v = buf[i:i + ln]
            i += ln
        elif wt == 5:
            v = buf[i:i + 4]
            i += 4
        else:
            raise ValueError('unsupported wire type %d' % wt)
        yield fno, wt, v


class UnigramModel:
    NORMAL, UNKNOWN, CONTROL, USER_DEFINED, UNUSED, BYTE = 1, 2, 3, 4, 5, 6

    def __init__(self, path_or_bytes):
        import numpy as np
        if isinstance(path_or_bytes, (bytes, bytearray)):
            buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, 'rb') as f:
                buf = f.read()
        self.pieces = []      # (str, np.float32, type)
        self.add_dummy_prefix = True
        self.remove_extra_whitespaces = True
        self.escape_whitespaces = True
        self.byte_fallback = False
        self.unk_id = 0
        self.model_type = 1
        for fno, wt, v in _pb_fields(buf):
            if fno == 1:
                piece, score, typ = '', np.float32(0.0), 1
                for f2, w2, v2 in _pb_fields(v):
                    if f2 == 1:
                        piece = bytes(v2).decode('utf-8')
                    elif f2 == 2:
                        score = np.float32(struct.unpack('<f', bytes(v2))[0])
                    elif f2 == 3:
                        typ = v2
                self.pieces.append((piece, score, typ))
            elif fno == 2:
                for f2, w2, v2 in _pb_fields(v):
                    if f2 == 3:
                        self.model_type = v2
                    elif f2 == 35:
                        self.byte_fallback = bool(v2)
                    elif f2 == 40:
                        self.unk_id = v2
            elif fno == 3:
                for f2, w2, v2 in _pb_fields(v):
                    if f2 == 3:
                        self.add_dummy_prefix = bool(v2)
                    elif f2 == 4:
                        self.remove_extra_whitespaces = bool(v2)
                    elif f2 == 5:
                        self.escape_whitespaces = bool(v2)
        self.piece_to_id = {}
        self.byte_to_id = {}
        self.max_len = 0
        mn = None
        for i, (p, s, t) in enumerate(self.pieces):
            if t in (self.NORMAL, self.USER_DEFINED, self.UNUSED):
                self.piece_to_id[p] = i
                self.max_len = max(self.max_len, len(p))
            if t == self.NORMAL:
                mn = s if mn is None or s < mn else mn
            if t == self.BYTE:
                self.byte_to_id[int(p[3:5], 16)] = i
            if t == self.UNKNOWN:
                self.unk_id = i
        self.min_score = mn if mn is not None else np.float32(0.0)
        self.unk_score = np.float32(self.min_score - np.float32(10.0))
        self.max_score = max((s for _, s, t in self.pieces if t == self.NORMAL), default=np.float32(0.0))

    def vocab_size(self):
        return len(self.pieces)


def spm_normalize(model, text):
    """identity charsmap + remove_extra_whitespaces + add_dummy_prefix + escape_whitespaces"""
    cps = [ord(c) for c in text]
    if model.remove_extra_whitespaces:
        out = []
        prev_space = True        # leading spaces are dropped
        for c in cps:
            if c == 0x20:
                if prev_space:
                    continue
                prev_space = True
            else:
                prev_space = False
            out.append(c)
        # trailing whitespace is removed from the ESCAPED output (sentencepiece normalizer.cc "Ignores tailing space"):
        # with escape_whitespaces a literal U+2581 at the end of the input goes too
        while out and (out[-1] == 0x20 or (model.escape_whitespaces and out[-1] == 0x2581)):
            out.pop()
        cps = out
    if not cps:
        return []
    if model.add_dummy_prefix:
        cps = [0x20] + cps
    if model.escape_whitespaces:
        cps = [0x2581 if c == 0x20 else c for c in cps]
    return cps


def unigram_encode(model, text):
    """ids of SentencePieceProcessor.EncodeAsIds(text) for a UNIGRAM model with byte_fallback"""
    import numpy as np
    cps = spm_normalize(model, text)
    n = len(cps)
    if n == 0:
        return []
    s = ''.join(map(chr, cps))
    NEG = None
    best = [NEG] * (n + 1)
    back = [(-1, -1)] * (n + 1)      # (start, id)
    best[0] = np.float32(0.0)
    p2i = model.piece_to_id
    pieces = model.pieces
    for i in range(n):
        if best[i] is NEG:
            continue
        bi = best[i]
        single = False
        for ln in range(1, min(model.max_len, n - i) + 1):
            pid = p2i.get(s[i:i + ln])
            if pid is None:
                continue
            _, sc, typ = pieces[pid]
            if typ == model.UNUSED:
                continue
            if typ == model.USER_DEFINED:
                sc = np.float32(np.float32(ln) * model.max_score - np.float32(0.1))
            # sentencepiece 0.2.1 unigram_model.cc (EncodeOptimized): `score` is the result of a ?: whose other arm is a
            # double expression, so the piece score is PROMOTED TO DOUBLE, the candidate is a double sum, it is compared
            # against the stored float as a double, and rounded to float only when stored
            cand = float(sc) + float(bi)
            t = i + ln
            if best[t] is NEG or cand > float(best[t]):
                best[t] = np.float32(cand)
                back[t] = (i, pid)
            if ln == 1:
                single = True
        if not single:
            cand = np.float32(bi + model.unk_score)
            t = i + 1
            if best[t] is NEG or cand > best[t]:
                best[t] = cand
                back[t] = (i, model.unk_id)
    ids_rev = []
    t = n
    while t > 0:
        i, pid = back[t]
        ids_rev.append((i, t, pid))
        t = i
    ids = []
    for i, t, pid in reversed(ids_rev):
        if pid == model.unk_id and model.byte_fallback:
            for b in s[i:t].encode('utf-8'):
                ids.append(model.byte_to_id[b])
        else:
            ids.append(pid)
    return ids


def _utf8_or_replacement(b):
    """a run of byte pieces: every well-formed UTF-8 sequence decodes, every other byte becomes U+FFFD"""
    out = []
    i = 0
    while i < len(b):
        for n in (1, 2, 3, 4):
            try:
                ch = bytes(b[i:i + n]).decode('utf-8')
            except UnicodeDecodeError:
                continue
            if len(ch) == 1:
                out.append(ch)
                i += n
                break
        else:
            out.append('\ufffd')
            i += 1
    return ''.join(out)


def unigram_decode(model, ids):
    """tokenizer.py:217-218 -> SentencePieceProcessor.DecodeIds (sentencepiece 0.2.1, sentencepiece_processor.cc Decode;
    behaviour pinned by tests/golden decode_fuzz): control pieces vanish, <unk> decodes to its surface ' \u2047 ', a run of
    byte pieces is decoded as UTF-8, U+2581 becomes a space, and while the text decoded so far is empty one leading U+2581 of
    a piece is consumed (add_dummy_prefix)"""
    out = []
    pend = []
    empty = True
    for i in ids:
        if i < 0 or i >= len(model.pieces):
            raise IndexError('piece id is out of range.')
        piece, _, typ = model.pieces[i]
        if typ == model.BYTE:
            pend.append(int(piece[3:5], 16))
            continue
        if pend:
            out.append(_utf8_or_replacement(pend))
            pend = []
            empty = False
        if typ == model.CONTROL:
            continue
        if typ == model.UNKNOWN:
            out.append(' \u2047 ')
            empty = False
            continue
        if empty and piece.startswith('\u2581'):
            piece = piece[1:]
        if piece:
            empty = False
        out.append(piece.replace('\u2581', ' '))
    if pend:
        out.append(_utf8_or_replacement(pend))
    return ''.join(out)


def unigram_pieces(model, text):
    return [model.pieces[i][0] for i in unigram_encode(model, text)]


# --------------------------------------------------------------------------------------------------
# ragged batch forms (the shapes the CUDA path returns; see include/akshar_b200.h)
# --------------------------------------------------------------------------------------------------
def utf8_len(cp):
    return 1 if cp < 0x80 else 2 if cp < 0x800 else 3 if cp < 0x10000 else 4


def cp_ends_to_byte_ends(cps, ends):
    """convert END indices in code points to END offsets in UTF-8 bytes (row-relative)"""
    pref = [0]
    for c in cps:
        pref.append(pref[-1] + utf8_len(c))
    return [pref[e] for e in ends]
