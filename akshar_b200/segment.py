"""Drop-in for the hot-path half of the reference's `akshar.segment` (src/akshar/segment.py:1-236), computed by the
CUDA grapheme-cluster / script-run kernel.  Same names, arguments, return types and label strings."""
import numpy as np

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


def analyze_text_composition_batch(texts, device=0):
    clusters, runs = engine(device).segment_batch(texts, clusters=True, runs=True)
    cs = clusters.splits.cpu().numpy()
    rs = runs.splits.cpu().numpy()
    ends = runs.values.cpu().numpy()
    tags = runs.extra.cpu().numpy()
    out = []
    for i, s in enumerate(texts):
        total = len(s)
        dev = rom = 0
        if total:
            b = s.encode('utf-8')
            prev = 0
            for e, t in zip(ends[rs[i]:rs[i + 1]].tolist(), tags[rs[i]:rs[i + 1]].tolist()):
                n = len(b[prev:e].decode('utf-8'))
                if t == 0:
                    dev += n
                elif t == 1:
                    rom += n
                prev = e
        out.append({
            'akshar_count': int(cs[i + 1] - cs[i]),
            'script_switches': int(rs[i + 1] - rs[i]) - 1,
            'devanagari_ratio': dev / total if total > 0 else 0,
            'roman_ratio': rom / total if total > 0 else 0,
        })
    return out


# ---- reference API (same signatures) ------------------------------------------------------------------
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
