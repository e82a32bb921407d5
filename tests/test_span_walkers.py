"""CPU build of the per-span walkers (akshar_b200/csrc/ak_text_core.cuh, compiled by g++ as a test aid) against the
oracle: cutting the buffer into arbitrary spans must never change the result -- the property the CUDA kernels rely on."""
import numpy as np
import pytest

import akshar_oracle as O
import oracle_batch as OB
import synth_corpus as sc
import walker_harness as W


import functools


@functools.lru_cache(None)
def _lines():
    lines = sc.adversarial(800, 21, 48) + sc.Corpus('social', 5).lines(8000) + sc.Corpus('hindi', 6).lines(5000)
    lines += ['', '', 'a', '', 'aaa', 'aa', '\n\n\n', '\u0958\u0958\u0958', 'e\u0301' * 5, '\u0928\u200c\u093c', '']
    return tuple(lines)


@functools.lru_cache(None)
def _exp_norm(flags):
    return OB.normalize_batch(_lines(), bool(flags & 1), bool(flags & 6))


@functools.lru_cache(None)
def _exp_seg(matras):
    return OB.segment_batch(_lines(), matras=matras)


@functools.lru_cache(None)
def _exp_runs():
    return OB.runs_batch(_lines())


@pytest.mark.parametrize('flags', [7, 6, 1, 0])
@pytest.mark.parametrize('span', [1, 3, 16, 64, 100000])
def test_normalize_spans(flags, span):
    lines = _lines()
    data, off = sc.pack(lines)
    exp, exp_off = _exp_norm(flags)
    out, out_off, st = W.normalize(data, off, flags=flags, span=span)
    assert st == 0
    assert np.array_equal(out_off, exp_off)
    assert out.tobytes() == exp.tobytes()


def test_normalize_random_spans():
    lines = _lines()
    data, off = sc.pack(lines)
    exp, exp_off = _exp_norm(7)
    rng = np.random.default_rng(3)
    for _ in range(3):
        out, out_off, st = W.normalize(data, off, flags=7, span=40, rng=rng)
        assert st == 0
        assert np.array_equal(out_off, exp_off)
        assert out.tobytes() == exp.tobytes()


@pytest.mark.parametrize('span', [1, 5, 32, 100000])
def test_segment_spans(span):
    lines = _lines()
    data, off = sc.pack(lines)
    ce, cs = _exp_seg(False)
    re_, rt, rs = _exp_runs()
    gce, gcs, gre, grt, grs, st = W.segment(data, off, flags=1 | 4, span=span)
    assert st == 0
    assert np.array_equal(gcs, cs) and np.array_equal(gce, ce)
    assert np.array_equal(grs, rs) and np.array_equal(gre, re_) and np.array_equal(grt, rt)
    me, ms = _exp_seg(True)
    gce, gcs, _, _, _, st = W.segment(data, off, flags=1 | 2, span=span)
    assert st == 0
    assert np.array_equal(gcs, ms) and np.array_equal(gce, me)


def test_bounded_lookback_flags_pathological():
    # a long run of Extend characters: a bounded backward walk must give up loudly, never answer wrongly
    lines = ['क' + 'ु' * 400 + 'ख']
    data, off = sc.pack(lines)
    _, _, _, _, _, st = W.segment(data, off, flags=1, span=16, limit=64)
    assert st & 4
    ce, cs = OB.segment_batch(lines)
    gce, gcs, _, _, _, st = W.segment(data, off, flags=1, span=16, limit=0)
    assert st == 0 and np.array_equal(gce, ce)


@pytest.mark.parametrize('flags,fn', [(1 | 8, O.semantic_normalize), (2 | 8, O.filter_garbage), (4 | 8, O.remove_elongations),
                                      (6 | 8, O.normalize_hinglish), (0, O.normalize_unicode)])
def test_single_stage_flags(flags, fn):
    lines = _lines()
    data, off = sc.pack(lines)
    exp = [fn(s).encode('utf-8') for s in lines]
    for span in (4, 32):
        out, out_off, st = W.normalize(data, off, flags=flags, span=span)
        assert st == 0
        b = out.tobytes()
        assert [b[out_off[i]:out_off[i + 1]] for i in range(len(lines))] == exp


@pytest.mark.parametrize('real', [30, 1, 2, 7])
def test_fast_lane_structure(real):
    """the fast kernel's chunk / halo / slow-lane structure gives the oracle's bytes and row offsets"""
    lines = list(_lines())
    data, off = sc.pack(lines)
    exp, exp_off = _exp_norm(7)
    out, out_off, st, n_slow = W.fast_normalize(data, off, real=real)
    assert st == 0
    assert np.array_equal(out_off, exp_off)
    assert out.tobytes() == exp.tobytes()


def test_fast_lane_is_mostly_fast():
    for kind in ('hinglish', 'hindi', 'social'):
        lines = sc.Corpus(kind, 8).lines(200000)
        data, off = sc.pack(lines)
        out, out_off, st, n_slow = W.fast_normalize(data, off)
        exp, exp_off = OB.normalize_batch(lines)
        assert st == 0 and np.array_equal(out_off, exp_off) and out.tobytes() == exp.tobytes()
        n_chunks = data.size / 16
        print(kind, 'slow chunks: %.2f %%' % (100.0 * n_slow / n_chunks))
        assert n_slow / n_chunks < 0.05


@pytest.mark.parametrize('real,stage_cap', [(30, 18), (1, 18), (4, 1)])
def test_segment_fast_structure(real, stage_cap):
    lines = list(_lines())
    data, off = sc.pack(lines)
    ce, cs = _exp_seg(False)
    re_, rt, rs = _exp_runs()
    gce, gcs, gre, grt, grs, st, n_slow = W.seg_fast(data, off, flags=1 | 4, real=real, stage_cap=stage_cap)
    assert st == 0
    assert np.array_equal(gcs, cs) and np.array_equal(gce, ce)
    assert np.array_equal(grs, rs) and np.array_equal(gre, re_) and np.array_equal(grt, rt)
    me, ms = _exp_seg(True)
    gce, gcs, _, _, _, st, _ = W.seg_fast(data, off, flags=1 | 2, real=real, stage_cap=stage_cap)
    assert st == 0
    assert np.array_equal(gcs, ms) and np.array_equal(gce, me)


def test_segment_fast_is_mostly_fast():
    for kind in ('hinglish', 'hindi', 'social'):
        lines = sc.Corpus(kind, 8).lines(200000)
        data, off = sc.pack(lines)
        gce, gcs, gre, grt, grs, st, n_slow = W.seg_fast(data, off, flags=5)
        ce, cs = OB.segment_batch(lines)
        re_, rt, rs = OB.runs_batch(lines)
        assert st == 0 and np.array_equal(gce, ce) and np.array_equal(gcs, cs)
        assert np.array_equal(gre, re_) and np.array_equal(grt, rt) and np.array_equal(grs, rs)
        print(kind, 'slow chunks: %.3f %%' % (100.0 * n_slow / (data.size / 16)))
        assert n_slow / (data.size / 16) < 0.02
