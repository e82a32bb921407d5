"""ctypes wrapper around tests/_build/libak_host_harness.so (CPU build of the span walkers; test aid only)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'tests', '_build', 'libak_host_harness.so')
_lib = None


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(ROOT, 'tests', 'csrc', 'host_harness.cpp')
    csrc = os.path.join(ROOT, 'akshar_b200', 'csrc')
    models = os.path.join(csrc, 'ak_models.cpp')
    deps = [src, models] + [os.path.join(csrc, f) for f in
                            ('ak_unicode.cuh', 'ak_bits.cuh', 'ak_norm3.cuh', 'ak_seg3.cuh', 'ak_bpe3.cuh', 'ak_text_core.cuh', 'ak_subword.cuh', 'ak_fast.cuh', 'ak_tok.cuh', 'ak_tok_host.h', 'ak_wordcache.cuh', 'ak_models.h', 'unicode_tables.inc')]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', SO, src, models])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(SO)
        _lib.hh_normalize.restype = ctypes.c_int64
        _lib.hh_signature.restype = ctypes.c_int64
        _lib.hh_fast_normalize3.restype = ctypes.c_int64
        _lib.hh_n3_roles.restype = ctypes.c_uint32
        _lib.hh_s3_roles.restype = ctypes.c_uint32
        _lib.hh_bpe.restype = ctypes.c_int64
        _lib.hh_tok.restype = ctypes.c_int64
        _lib.hh_unigram.restype = ctypes.c_int64
        _lib.hh_error.restype = ctypes.c_char_p
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_spans(off, span, rng=None):
    """span boundaries from off[0] to off[-1]+1; span=int fixed size, or random sizes in [1, span] with rng"""
    lo, hi = int(off[0]), int(off[-1]) + 1
    b = [lo]
    while b[-1] < hi:
        step = span if rng is None else int(rng.integers(1, span + 1))
        b.append(min(hi, b[-1] + step))
    return np.array(b, dtype=np.int64)


def normalize(data, off, flags=7, span=32, limit=0, rng=None):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    spans = make_spans(off, span, rng)
    out = np.zeros(int(data.size) * 3 + 16, dtype=np.uint8)
    out_off = np.full(off.size, -1, dtype=np.int64)
    st = ctypes.c_uint32(0)
    n = lib().hh_normalize(_p(data), _p(off), ctypes.c_int64(off.size - 1), ctypes.c_uint32(flags), _p(spans),
                           ctypes.c_int64(spans.size - 1), ctypes.c_int64(limit), _p(out), _p(out_off), ctypes.byref(st))
    return out[:n], out_off, st.value


def segment(data, off, flags=1, span=32, limit=0, rng=None):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    spans = make_spans(off, span, rng)
    cap = int(data.size) + 1
    ce = np.zeros(cap, dtype=np.int32)
    cs = np.full(off.size, -1, dtype=np.int64)
    re_ = np.zeros(cap, dtype=np.int32)
    rt = np.zeros(cap, dtype=np.uint8)
    rs = np.full(off.size, -1, dtype=np.int64)
    tot = np.zeros(2, dtype=np.int64)
    st = ctypes.c_uint32(0)
    lib().hh_segment(_p(data), _p(off), ctypes.c_int64(off.size - 1), ctypes.c_uint32(flags), _p(spans),
                     ctypes.c_int64(spans.size - 1), ctypes.c_int64(limit), _p(ce), _p(cs), _p(re_), _p(rt), _p(rs),
                     ctypes.c_int64(cap), _p(tot), ctypes.byref(st))
    return ce[:tot[0]], cs, re_[:tot[1]], rt[:tot[1]], rs, st.value


def signature(data, off):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    out = np.zeros(int(data.size) * 2 + 16, dtype=np.uint8)
    out_off = np.zeros(off.size, dtype=np.int64)
    n = lib().hh_signature(_p(data), _p(off), ctypes.c_int64(off.size - 1), _p(out), _p(out_off))
    return out[:n], out_off


def load_bpe(path):
    b = open(path, 'rb').read()
    if lib().hh_load_bpe(b, ctypes.c_int64(len(b))) != 0:
        raise ValueError(lib().hh_error().decode())
    return lib().hh_bpe_vocab_size()


def load_spm(path):
    b = open(path, 'rb').read()
    if lib().hh_load_spm(b, ctypes.c_int64(len(b))) != 0:
        raise ValueError(lib().hh_error().decode())
    return lib().hh_spm_vocab_size()


def bpe(data, off, span=32, limit=0, rng=None, stage_cap=40):
    """the kernel's three-pass protocol: encode; if NFC would change the text, NFC it and encode the copy"""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    for attempt in range(2):
        spans = make_spans(off, span, rng)
        cap = int(data.size) + 2 * off.size + 16
        ids = np.zeros(cap, dtype=np.int32)
        splits = np.full(off.size, -1, dtype=np.int64)
        ch = ctypes.c_int(0)
        st = ctypes.c_uint32(0)
        n = lib().hh_bpe(_p(data), _p(off), ctypes.c_int64(off.size - 1), _p(spans), ctypes.c_int64(spans.size - 1),
                         ctypes.c_int64(limit), ctypes.c_int(stage_cap), _p(ids), ctypes.c_int64(cap), _p(splits),
                         ctypes.byref(ch), ctypes.byref(st))
        if not ch.value:
            return ids[:n], splits, st.value, attempt
        assert attempt == 0, 'NFC copy still not normalized'
        data, off, _ = normalize(data, off, flags=0, span=span, limit=limit, rng=rng)
        data = np.ascontiguousarray(data)
    raise AssertionError


def unigram(data, off):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    cap = int(data.size) * 4 + 2 * off.size + 16
    ids = np.zeros(cap, dtype=np.int32)
    splits = np.zeros(off.size, dtype=np.int64)
    n = lib().hh_unigram(_p(data), _p(off), ctypes.c_int64(off.size - 1), _p(ids), ctypes.c_int64(cap), _p(splits))
    return ids[:n], splits


def fast_normalize3(data, off, real=30, flags=7):
    """normalize_text (default flags) through the bit-parallel kernel's lane / exchange / slow-lane structure"""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    out = np.zeros(int(data.size) * 3 + 64, dtype=np.uint8)
    out_off = np.full(off.size, -1, dtype=np.int64)
    st = ctypes.c_uint32(0)
    ns = ctypes.c_int64(0)
    n = lib().hh_fast_normalize3(_p(data), _p(off), ctypes.c_int64(off.size - 1), ctypes.c_int(real), ctypes.c_uint32(flags), _p(out), _p(out_off),
                                 ctypes.byref(st), ctypes.byref(ns))
    return out[:n], out_off, st.value, ns.value


def n3_roles(byte):
    return int(lib().hh_n3_roles(ctypes.c_uint32(byte)))


def n3_planes(b32):
    b = np.ascontiguousarray(b32, dtype=np.uint8)
    P = np.zeros(8, dtype=np.uint32)
    lib().hh_n3_planes(_p(b), _p(P))
    return P


def s3_roles(byte):
    return int(lib().hh_s3_roles(ctypes.c_uint32(byte)))


def seg_fast3(data, off, flags=1, real=30):
    """clusters / script runs through the bit-parallel kernel's lane structure (ak_seg3.cuh)"""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    cap = int(data.size) + off.size + 1
    ce = np.zeros(cap, dtype=np.int32)
    cs = np.full(off.size, -1, dtype=np.int64)
    re_ = np.zeros(cap, dtype=np.int32)
    rt = np.zeros(cap, dtype=np.uint8)
    rs = np.full(off.size, -1, dtype=np.int64)
    tot = np.zeros(2, dtype=np.int64)
    st = ctypes.c_uint32(0)
    ns = ctypes.c_int64(0)
    lib().hh_seg_fast3(_p(data), _p(off), ctypes.c_int64(off.size - 1), ctypes.c_uint32(flags), ctypes.c_int(real),
                       _p(ce), _p(cs), _p(re_), _p(rt), _p(rs), ctypes.c_int64(cap), _p(tot), ctypes.byref(st), ctypes.byref(ns))
    return ce[:tot[0]], cs, re_[:tot[1]], rt[:tot[1]], rs, st.value, ns.value


def lines(data, span=32):
    """file bytes -> rows (ak_lines.cuh) on the CPU -> list[bytes]"""
    data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data, dtype=np.uint8)
    pad = np.concatenate([data, np.zeros(8, dtype=np.uint8)])
    cap = int(data.size) + 2
    b = np.zeros(cap, dtype=np.int64)
    e = np.zeros(cap, dtype=np.int64)
    st = ctypes.c_uint32(0)
    lib().hh_lines.restype = ctypes.c_int64
    n = lib().hh_lines(_p(pad), ctypes.c_int64(data.size), ctypes.c_int(span), _p(b), _p(e), ctypes.c_int64(cap), ctypes.byref(st))
    assert st.value == 0
    raw = data.tobytes()
    return [raw[b[i]:e[i]] for i in range(n)]


def lines32(data, lead=0):
    """file bytes -> rows through the kernels' 32-byte mask lanes (akln_*) -> list[bytes]"""
    data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data, dtype=np.uint8)
    pad = np.concatenate([data, np.zeros(8, dtype=np.uint8)])
    cap = int(data.size) + 2
    b = np.zeros(cap, dtype=np.int64)
    e = np.zeros(cap, dtype=np.int64)
    st = ctypes.c_uint32(0)
    lib().hh_lines32.restype = ctypes.c_int64
    n = lib().hh_lines32(_p(pad), ctypes.c_int64(data.size), ctypes.c_int(lead), _p(b), _p(e), ctypes.c_int64(cap), ctypes.byref(st))
    assert st.value == 0
    raw = data.tobytes()
    return [raw[b[i]:e[i]] for i in range(n)]


def wordtok(data, off, mode=0, real=30):
    """the word tokenizers (ak_wordtok.cuh) through the kernel's lane structure -> (begin, end, splits, row_flags, status)"""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    cap = int(data.size) + 8
    wb = np.full(cap, -1, dtype=np.int32)
    we = np.full(cap, -1, dtype=np.int32)
    sp = np.full(off.size, -1, dtype=np.int64)
    fl = np.zeros(max(off.size - 1, 1), dtype=np.uint8)
    st = ctypes.c_uint32(0)
    lib().hh_wordtok.restype = ctypes.c_int64
    n = lib().hh_wordtok(_p(data), _p(off), ctypes.c_int64(off.size - 1), ctypes.c_int(mode), ctypes.c_int(real), _p(wb), _p(we),
                         ctypes.c_int64(cap), _p(sp), _p(fl), ctypes.byref(st))
    return wb[:n], we[:n], sp, fl[:off.size - 1], st.value


def decode(kind, form, rows):
    """ids -> text (ak_decode.cuh) on the CPU; rows = list of id lists -> (list[bytes], status)"""
    n = sum(len(r) for r in rows)
    ids = np.array([i for r in rows for i in r], dtype=np.int32) if n else np.zeros(1, dtype=np.int32)
    sp = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in rows], out=sp[1:])
    cap = 64 * n + 16
    out = np.zeros(cap, dtype=np.uint8)
    off = np.zeros(len(rows) + 1, dtype=np.int64)
    st = ctypes.c_uint32(0)
    lib().hh_decode.restype = ctypes.c_int64
    tot = lib().hh_decode(ctypes.c_int(kind), ctypes.c_int(form), _p(ids), ctypes.c_int64(n), _p(sp), ctypes.c_int64(len(rows)), _p(out),
                          ctypes.c_int64(cap), _p(off), ctypes.byref(st))
    b = out[:tot].tobytes()
    return [b[off[i]:off[i + 1]] for i in range(len(rows))], st.value


def tok(kind, data, off, real=30, cache_bits=14, prewarm=1, u16=False, splits_i32=False, cap=None):
    """the event-stream encoders (ak_tok.cuh) on the CPU: lanes -> event slots -> row fix -> resolve -> check -> emit.
    kind 0 BPE, 1 Unigram; cap = event slots per emulated warp tile -> (ids, splits, status, stats dict)"""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    id_cap = int(data.size) * 4 + 2 * off.size + 16
    ids = np.zeros(id_cap, dtype=np.uint16 if u16 else np.int32)
    splits = np.full(off.size, -1, dtype=np.int32 if splits_i32 else np.int64)
    st = ctypes.c_uint32(0)
    stats = np.zeros(5, dtype=np.int64)
    if cap is None:
        cap = 64 * real + 8             # a row start and a word start at every byte
    n = lib().hh_tok(ctypes.c_int(kind), _p(data), _p(off), ctypes.c_int64(off.size - 1), ctypes.c_int(real), ctypes.c_int(cap), ctypes.c_int(cache_bits),
                     ctypes.c_int(prewarm), ctypes.c_int(1 if u16 else 0), ctypes.c_int(1 if splits_i32 else 0), _p(ids),
                     ctypes.c_int64(id_cap), _p(splits), ctypes.byref(st), _p(stats))
    assert n >= 0, 'the model does not have the shape the word-wise Unigram path needs'
    return ids[:n], splits, st.value, dict(events=int(stats[0]), flagged_rows=int(stats[1]), exact_words=int(stats[2]), misses=int(stats[3]), slots_needed=int(stats[4]))
