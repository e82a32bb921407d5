"""Drop-in for the hot-path half of the reference's `akshar.segment` (src/akshar/segment.py:1-236), computed by the
CUDA grapheme-cluster / script-run kernel.  Same names, arguments, return types and label strings."""
import numpy as np

from . import _lib as C
from .batch import engine

MATRA_RANGES = [          # reference segment.py:20-24
    (0x093E, 0x094C),
    (0x0900, 0x0902),
    (0x0951, 0x0954),
]
_TAGS = ['devanagari', 'roman', 'digit', 'punct', 'other']


class _GpuGraphemePattern:
    """stands in for the compiled `regex` pattern object `akshar_PAT` (segment.py:14): findall == \\X clusters"""
    pattern = r'\X'

    def findall(self, text):
        return segment_akshars(text)


akshar_PAT = _GpuGraphemePattern()


def is_matra(char):
    """reference segment.py:26-37"""
    if not char:
        return False
    cp = ord(char[0])
    return any(lo <= cp <= hi for lo, hi in MATRA_RANGES)


def _slices(texts, ragged):
    """cut every row at its END byte offsets -> list[list[str]]"""
    ends = ragged.values.cpu().numpy()
    splits = ragged.splits.cpu().numpy()
    out = []
    for i, s in enumerate(texts):
        b = s.encode('utf-8')
        e = ends[splits[i]:splits[i + 1]]
        prev = 0
        parts = []
        for x in e.tolist():
            parts.append(b[prev:x].decode('utf-8'))
            prev = x
        out.append(parts)
    return out


# ---- batch entry points (new) -----------------------------------------------------------------------
def segment_akshars_batch(texts, matras=False, as_device=False, device=0):
    """-> list[list[str]]; with as_device=True the Ragged(int32 END byte offsets per row, int64 row_splits)"""
    clusters, _ = engine(device).segment_batch(texts, clusters=True, matras=matras, runs=False)
    return clusters if as_device else _slices(texts, clusters)


def normalize_and_segment_batch(texts, normalize_roman=True, clean_hinglish=True, matras=False, device=0):
    """normalize_text then segment_akshars of a batch in ONE library call (akshar_normalize_segment_batch) and two
    host synchronisations: -> (list of normalized strings, list of akshar lists).  What `aksharTokenizer.tokenize` without a
    model is; the boundaries come back as one bit per byte."""
    import numpy as np
    import torch
    eng = engine(device)
    norm, mk = eng.normalize_segment_batch(texts, normalize_roman, clean_hinglish, clusters=True, matras=matras, runs=False)
    n, rows = norm.end, norm.n_rows
    W = (n + 32) // 32
    # everything that goes back in one pinned buffer: text, row offsets, mask words
    pin = eng.__dict__.get('_small_pin')
    need = n + 8 * (rows + 1) + 4 * W + 64
    if pin is None or pin.numel() < need:
        pin = torch.empty(max(need, 1 << 16), dtype=torch.uint8).pin_memory()
        eng._small_pin = pin
    a = (n + 7) & ~7
    b = a + 8 * (rows + 1)
    pin[:n].copy_(norm.data[:n], non_blocking=True)
    pin[a:b].view(torch.int64).copy_(norm.offsets, non_blocking=True)
    pin[b:b + 4 * W].view(torch.int32).copy_(mk['cluster'][:W], non_blocking=True)
    torch.cuda.current_stream(eng.device).synchronize()
    buf = pin.numpy()
    data = buf[:n].tobytes()
    off = buf[a:b].view(np.int64)
    bits = np.unpackbits(buf[b:b + 4 * W], bitorder='little')
    ends = np.flatnonzero(bits)                      # byte positions where a cluster ends, ascending
    cut = np.searchsorted(ends, off, side='right')   # ends at or before each row start
    strings, akshars = [], []
    for i in range(rows):
        lo, hi = int(off[i]), int(off[i + 1])
        strings.append(data[lo:hi].decode('utf-8'))
        prev = lo
        parts = []
        for e in ends[cut[i]:cut[i + 1]].tolist():
            parts.append(data[prev:e].decode('utf-8'))
            prev = e
        akshars.append(parts)
    return strings, akshars


def detect_code_switches_batch(texts, as_device=False, device=0):
    """-> list[list[(segment, label)]]; as_device=True: Ragged(run END offsets, row_splits, uint8 tags)"""
    _, runs = engine(device).segment_batch(texts, clusters=False, runs=True)
    if as_device:
        return runs
    segs = _slices(texts, runs)
    tags = runs.extra.cpu().numpy()
    splits = runs.splits.cpu().numpy()
    out = []
    for i, parts in enumerate(segs):
        tg = tags[splits[i]:splits[i + 1]].tolist()
        out.append([(p, None if t == 255 else _TAGS[t]) for p, t in zip(parts, tg)])
    return out


def analyze_text_composition_batch(texts, as_device=False, device=0):
    """analyze_text_composition over a batch (reference segment.py:210-236): the five counts of every row come off the
    device (Engine.composition_batch); as_device=True returns them as an int32 [n, 5] tensor
    (akshars, script runs, code points, code points in devanagari runs, in roman runs)"""
    stats = engine(device).composition_batch(texts)
    if as_device:
        return stats
    out = []
    for ak, nr, total, dev, rom in stats.cpu().numpy().tolist():
        out.append({
            'akshar_count': ak,
            'script_switches': nr - 1,
            'devanagari_ratio': dev / total if total > 0 else 0,
            'roman_ratio': rom / total if total > 0 else 0,
        })
    return out


def _merged(texts, rule, device):
    eng = engine(device)
    b = eng.put(texts)
    clusters, _ = eng.segment_batch(b, clusters=True, matras=False, runs=False)
    return _slices(texts, eng.merge_clusters_batch(b, clusters, rule))


def akshara_level_tokenization_batch(texts, device=0):
    """features.py:28-55 over a batch: consecutive clusters that hold a halant are one akshara"""
    return _merged(texts, C.MERGE_AKSHARA, device)


def preserve_nukta_batch(texts, device=0):
    """features.py:173-206 over a batch: a cluster that holds a nukta takes the next cluster with it"""
    return _merged(texts, C.MERGE_NUKTA, device)


def akshara_level_tokenization(text):
    """reference features.py:28-55"""
    return akshara_level_tokenization_batch([text])[0]


def preserve_nukta(text):
    """reference features.py:173-206"""
    return preserve_nukta_batch([text])[0]


def _word_slices(tb, begin, end, splits, rows=None):
    """cut the rows of the device TextBatch `tb` at the token offsets -> list[list[str]] (all rows, or just `rows`)"""
    n = tb.n_rows
    data = tb.data[:max(tb.end, 0)].cpu().numpy().tobytes() if tb.end > 0 else b''
    off = tb.offsets.cpu().numpy()
    wb, we, sp = begin.cpu().numpy(), end.cpu().numpy(), splits.cpu().numpy()
    out = []
    for i in (range(n) if rows is None else rows):
        base = int(off[i])
        lo, hi = int(sp[i]), int(sp[i + 1])
        out.append([data[base + b:base + e].decode('utf-8') for b, e in zip(wb[lo:hi].tolist(), we[lo:hi].tolist())])
    return out


def word_tokenize_hindi_batch(texts, use_morphology=False, device=0):
    """word_tokenize_hindi over a batch: normalize_text on the device, then the word kernel over the normalized text"""
    eng = engine(device)
    norm = eng.normalize_batch(texts)
    wb, we, sp = eng.word_tokenize_batch(norm, rule=C.WORDS_HINDI)
    return _word_slices(norm, wb, we, sp)


word_tokenize_sanskrit_batch = word_tokenize_hindi_batch      # the reference's two loops are the same (segment.py:335-362)


def word_tokenize_batch(texts, language='auto', use_morphology=False, device=0):
    """word_tokenize over a batch (reference segment.py:365-401)"""
    lang = language.lower()
    if language != 'auto':
        if lang in ('hindi', 'hi', 'hin', 'sanskrit', 'sa', 'san', 'skr'):
            return word_tokenize_hindi_batch(texts, use_morphology, device)
    eng = engine(device)
    raw = eng.put(texts)
    wb, we, sp, fl = eng.word_tokenize_batch(raw, rule=C.WORDS_SPLIT, row_flags=True)
    out = _word_slices(raw, wb, we, sp)
    if language != 'auto':
        return out                                            # unknown language: whitespace split (segment.py:399-401)
    deva = [i for i, f in enumerate(fl.cpu().numpy().tolist()) if f]
    if deva:
        # rows with a code point of U+0900-097F take the Hindi route (segment.py:384-388)
        sub = word_tokenize_hindi_batch([texts[i] for i in deva], use_morphology, device)
        for i, w in zip(deva, sub):
            out[i] = w
    return out


# ---- reference API (same signatures) ------------------------------------------------------------------
def word_tokenize_hindi(text, use_morphology=False):
    """reference segment.py:239-300.  `use_morphology` asks the reference for its Morfessor model when one is loaded and
    silently takes this same loop when none is (morph.py is out of scope here: no model is ever loaded)."""
    return word_tokenize_hindi_batch([text], use_morphology)[0]


def word_tokenize_sanskrit(text, use_morphology=False):
    """reference segment.py:303-362"""
    return word_tokenize_sanskrit_batch([text], use_morphology)[0]


def word_tokenize(text, language='auto', use_morphology=False):
    """reference segment.py:365-401"""
    return word_tokenize_batch([text], language, use_morphology)[0]


def segment_akshars(text, matras=False, separate_matras=None):
    """reference segment.py:40-125"""
    if separate_matras is not None:
        matras = separate_matras
    return segment_akshars_batch([text], matras=bool(matras))[0]


_PUNCT = " .,!?;:'\"()-[]{}"      # the literal of reference segment.py:141


def identify_script(char):
    """reference segment.py:128-147 (single character -> label)"""
    runs = engine().segment_batch([char], clusters=False, runs=True)[1]
    t = int(runs.extra.cpu().numpy()[0]) if runs.values.numel() else 255
    if t != 255:
        return _TAGS[t]
    # a lone punct / digit character forms a run without a label; the two are told apart by the reference's literal
    return 'punct' if char in _PUNCT else 'digit'


def detect_code_switches(text):
    """reference segment.py:150-201"""
    return detect_code_switches_batch([text])[0]


def segment_by_script(text):
    """reference segment.py:204-207"""
    return [seg for seg, _ in detect_code_switches(text)]


def analyze_text_composition(text):
    """reference segment.py:210-236"""
    return analyze_text_composition_batch([text])[0]
