// Per-span walkers for the text stages of the hot path.
//
// The concatenated UTF-8 buffer is cut into fixed-size byte spans; ONE thread walks one span.  A walker owns
// every code point whose lead byte lies in its span.  Everything a decision needs from before the span (NFC
// segment membership, elongation-run state, grapheme-rule state, current script) is RECONSTRUCTED by reading
// backwards from the span start to the nearest synchronising point (row start, NFC segment head, a base
// character, a strong-script character).  Spans are therefore independent: no carried state between threads,
// tiles or CTAs, and the only cross-thread quantity is the exclusive prefix of output counts.
//
// Backward walks are bounded by `limit` bytes; exceeding it sets AK_ST_PATHOLOGICAL and the host re-runs the
// batch with one span per row (limit = 0 -> unlimited), which needs no reconstruction at all.
#pragma once
#include "ak_unicode.cuh"

#define AK_MAXSEG 256          // decomposed code points per NFC segment handled by the slow lane
enum {
    AK_ST_OVERFLOW = 1,        // an output buffer was too small; totals are still exact
    AK_ST_NFC_SEGMENT = 2,     // a non-inert NFC segment exceeded AK_MAXSEG decomposed code points
    AK_ST_PATHOLOGICAL = 4,    // a bounded look-back gave up; result invalid, re-run row-sequentially
    AK_ST_ALPHABET = 8,        // subword encode met a code point outside the supported alphabet
    AK_ST_SPIN = 16,           // decoupled look-back exceeded its spin budget (never expected)
    AK_ST_WORD = 32,           // BPE word longer than the per-word symbol capacity
    AK_ST_INTERNAL = 64,       // an internal consistency check failed (never expected)
    AK_ST_BAD_ID = 128,        // decode: a token id outside the vocabulary (SentencePiece: 'piece id is out of range.')
};
#define AK_NORM_ROMAN 1u       // semantic_normalize: reference normalize_text(normalize_roman=True)
#define AK_NORM_FILTER 2u      // filter_garbage      \ together: normalize_hinglish =
#define AK_NORM_COLLAPSE 4u    // remove_elongations  /  reference normalize_text(clean_hinglish=True)
#define AK_NORM_CLEAN 6u
#define AK_NORM_NO_NFC 8u      // skip normalize_unicode (the stand-alone stage functions of normalize.py)
#define AK_SEG_CLUSTERS 1u
#define AK_SEG_MATRAS 2u       // reference segment_akshars(matras=True)
#define AK_SEG_RUNS 4u

// first r in [lo, hi] with off[r] >= p  (off[hi] >= p is guaranteed by the caller)
AK_HD int64_t ak_row_lower_bound(const int64_t* off, int64_t lo, int64_t hi, int64_t p) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (off[mid] >= p) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// =================================================================================================
// normalize  (reference normalize.py:117-148 = NFC -> Roman lowercase -> allow-list -> collapse runs >= 3)
// =================================================================================================
struct AkNormSink {
    uint8_t* out;      // nullptr: count only
    int64_t cnt;
    int64_t cap;       // bytes `out` can take; the count stays exact beyond it
};
AK_HD void ak_sink_put(AkNormSink& s, uint32_t cp) {
    const int len = ak_utf8_len(cp);
    if (s.out && s.cnt + len <= s.cap) ak_encode(cp, s.out + s.cnt);
    s.cnt += len;
}

struct AkCollapse {
    uint32_t c;        // last kept code point of the row, 0xFFFFFFFF if none
    int n;             // min(length of the run of c ending here, 3); '\n' always 1
    bool pending;      // the span's own 2nd element of a run awaits the next kept code point (emit iff it differs)
};

// one code point of the NFC stream through lowercase + filter; returns how many code points survive (0..2)
AK_HD int ak_post_nfc(const AkTables& T, uint32_t cp, uint32_t props, uint32_t flags, uint32_t* o) {
    int n = 1;
    o[0] = cp;
    if ((flags & AK_NORM_ROMAN) && AK_LATIN_LOWER(props)) {
        o[0] = ak_latin_lower(T, cp);
        if (cp == 0x130u) { o[1] = 0x307u; n = 2; }
        if (flags & AK_NORM_FILTER) props = ak_props(T, o[0]);
    }
    if (flags & AK_NORM_FILTER) {
        int m = 0;
        if (AK_ALLOW(props)) o[m++] = o[0];
        if (n == 2 && AK_ALLOW(ak_props(T, o[1]))) o[m++] = o[1];
        n = m;
    }
    return n;
}

// Elongation collapse, reference normalize.py:56 `(.)\1{2,}` -> `\1`: in a maximal run of equal kept code points
// (never '\n') the 1st is emitted, the 2nd iff the run stops there, the rest never.  Every element is emitted by
// the span that owns it: the 2nd of a run is held back until the next kept code point is known (look-ahead past the
// span end when necessary), nothing is ever emitted on behalf of another span.
AK_HD void ak_collapse_feed(AkCollapse& st, uint32_t y, AkNormSink& sink) {
    if (y == st.c && y != 0x0Au) {
        if (st.n == 1) { st.n = 2; st.pending = true; }
        else { st.n = 3; st.pending = false; }
    } else {
        if (st.pending) { ak_sink_put(sink, st.c); st.pending = false; }
        ak_sink_put(sink, y);
        st.c = y;
        st.n = 1;
    }
}
// the run is over (row end: next == 0xFFFFFFFF) or the span is: resolve the held-back element
AK_HD void ak_collapse_close(AkCollapse& st, uint32_t next, AkNormSink& sink) {
    if (st.pending) {
        if (next != st.c) ak_sink_put(sink, st.c);
        st.pending = false;
    }
}

// scan the NFC segment that starts at p (p is a head, or the first code point of a row): returns its end and
// whether NFC can change it ("troubled").  The fast lane of the kernels uses the SAME definition code point by code
// point, so both agree on which segments are emitted whole by the owner of their head.  `scan_limit` bounds the scan.
AK_HD int64_t ak_scan_segment(const AkTables& T, const uint8_t* t, int64_t p, int64_t re, bool& trouble,
                              int64_t scan_limit, uint32_t& status) {
    int len;
    uint32_t cp = ak_decode(t, p, re, len);
    uint32_t w = ak_props(T, cp);
    trouble = AK_QC(w) != 0;
    uint32_t prev_ccc = AK_CCC(w);
    // a QC=Maybe mark directly after an atomic starter that composes with nothing cannot change (ak_unicode.cuh)
    bool after_inert_base = AK_INERT_BASE(w);
    int64_t q = p + len;
    while (q < re) {
        cp = ak_decode(t, q, re, len);
        w = ak_props(T, cp);
        if (AK_NFC_HEAD(w)) break;
        uint32_t cc = AK_CCC(w);
        uint32_t qc = AK_QC(w);
        if (qc == 1u || (qc == 2u && !after_inert_base) || (cc != 0 && prev_ccc > cc)) trouble = true;
        after_inert_base = false;
        prev_ccc = cc;
        q += len;
        if (scan_limit > 0 && q - p > scan_limit) { status |= AK_ST_PATHOLOGICAL; break; }
    }
    return q;
}

// NFC of the segment [h, hend) into buf; returns count
AK_HD_NOINLINE int ak_nfc_segment(const AkTables& T, const uint8_t* t, int64_t h, int64_t hend, uint32_t* buf,
                                  uint32_t& status) {
    int n = 0;
    int64_t q = h;
    while (q < hend) {
        int len;
        uint32_t cp = ak_decode(t, q, hend, len);
        if (n + 4 > AK_MAXSEG) { status |= AK_ST_NFC_SEGMENT; break; }
        n = ak_decompose(T, cp, ak_props(T, cp), buf, n);
        q += len;
    }
    return ak_nfc_inplace(T, buf, n);
}

// head of the NFC segment containing the code point that starts at q (rs <= q)
AK_HD int64_t ak_find_head(const AkTables& T, const uint8_t* t, int64_t q, int64_t rs, int64_t re, int64_t limit,
                           uint32_t& status) {
    int64_t q0 = q;
    while (q > rs) {
        int len;
        uint32_t cp = ak_decode(t, q, re, len);
        if (AK_NFC_HEAD(ak_props(T, cp))) break;
        q = ak_prev_start(t, q, rs);
        if (limit > 0 && q0 - q > limit) { status |= AK_ST_PATHOLOGICAL; break; }
    }
    return q;
}

// last (up to 3) kept code points of the row before position p, newest first.  p is a segment boundary or an
// inert follower (see ak_norm_span).
AK_HD_NOINLINE int ak_prev_kept(const AkTables& T, const uint8_t* t, int64_t p, int64_t rs, int64_t re, uint32_t flags,
                                int64_t limit, uint32_t* k, uint32_t& status) {
    int nk = 0;
    int64_t q = p;
    while (nk < 3 && q > rs) {
        if (limit > 0 && p - q > limit) { status |= AK_ST_PATHOLOGICAL; break; }
        int64_t last = ak_prev_start(t, q, rs);
        int64_t h = last;
        bool trouble = false;
        if (!(flags & AK_NORM_NO_NFC)) {
            h = ak_find_head(T, t, last, rs, re, limit, status);
            // whole-segment verdict (on the first step the segment may extend past p, but only when it is inert)
            ak_scan_segment(T, t, h, re, trouble, limit, status);
        }
        if (!trouble) {
            // inert: code points map one by one; walk them backwards
            int64_t c = q;
            while (c > h && nk < 3) {
                int64_t cs = ak_prev_start(t, c, rs);
                int len;
                uint32_t cp = ak_decode(t, cs, re, len);
                uint32_t o[2];
                int m = ak_post_nfc(T, cp, ak_props(T, cp), flags, o);
                for (int i = m - 1; i >= 0 && nk < 3; --i) k[nk++] = o[i];
                c = cs;
            }
        } else {
            uint32_t buf[AK_MAXSEG];
            int n = ak_nfc_segment(T, t, h, q, buf, status);
            for (int i = n - 1; i >= 0 && nk < 3; --i) {
                uint32_t o[2];
                int m = ak_post_nfc(T, buf[i], ak_props(T, buf[i]), flags, o);
                for (int j = m - 1; j >= 0 && nk < 3; --j) k[nk++] = o[j];
            }
        }
        q = h;
    }
    return nk;
}

// first kept code point produced by the text from q (a code-point start inside row [rs, re)) on, or 0xFFFFFFFF
// when the row ends first.  q is either a segment start or inside an inert segment (a troubled one is always
// consumed whole by the main loop).
AK_HD_NOINLINE uint32_t ak_next_kept(const AkTables& T, const uint8_t* t, int64_t q, int64_t rs, int64_t re, uint32_t flags,
                                     int64_t limit, uint32_t& status) {
    const int64_t q0 = q;
    while (q < re) {
        if (limit > 0 && q - q0 > limit) { status |= AK_ST_PATHOLOGICAL; break; }
        int len;
        uint32_t cp = ak_decode(t, q, re, len);
        uint32_t w = ak_props(T, cp);
        if (!(flags & AK_NORM_NO_NFC) && (q == rs || AK_NFC_HEAD(w))) {
            bool trouble;
            int64_t hend = ak_scan_segment(T, t, q, re, trouble, limit, status);
            if (trouble) {
                uint32_t buf[AK_MAXSEG];
                int n = ak_nfc_segment(T, t, q, hend, buf, status);
                for (int i = 0; i < n; ++i) {
                    uint32_t o[2];
                    if (ak_post_nfc(T, buf[i], ak_props(T, buf[i]), flags, o) > 0) return o[0];
                }
                q = hend;
                continue;
            }
        }
        uint32_t o[2];
        if (ak_post_nfc(T, cp, w, flags, o) > 0) return o[0];
        q += len;
    }
    return 0xFFFFFFFFu;
}

// Walk span [s, e).  `off` = absolute row offsets (n_rows + 1 entries), search window [r_lo, r_hi] must satisfy
// off[r_lo] <= first owned position or r_lo == 0, and off[r_hi] >= e or r_hi == n_rows.
// out: nullptr for the counting pass, else the address where THIS span's first output byte goes.
// out_off: nullptr or the output row-offset array; entry r gets (out_base + bytes emitted before row r).
AK_HD_NOINLINE int64_t ak_norm_span(const AkTables& T, const uint8_t* t, const int64_t* off, int64_t n_rows,
                                    int64_t r_lo, int64_t r_hi, int64_t s, int64_t e, uint32_t flags, int64_t limit,
                                    uint8_t* out, int64_t* out_off, int64_t out_base, uint32_t& status,
                                    int64_t out_cap = 0x7FFFFFFFFFFFFFFFll) {
    const int64_t total_end = off[n_rows];
    const bool clean = (flags & AK_NORM_COLLAPSE) != 0;
    const bool nfc = (flags & AK_NORM_NO_NFC) == 0;
    AkNormSink sink;
    sink.out = out;
    sink.cnt = 0;
    sink.cap = out_cap;
    int64_t p = s;
    // skip continuation bytes of a code point owned by the previous span (never past a row start: valid UTF-8)
    if (p < total_end && p > off[0]) {
        // (only inside a row: a row that BEGINS with continuation bytes -- bytes that are not UTF-8 -- keeps them, so that
        // its row-start event is still met; nothing is skipped across the next row start either)
        const int64_t nr0 = ak_row_lower_bound(off, r_lo, r_hi, p);
        int k = 0;
        while (off[nr0] != s && p < e && p < total_end && p < off[nr0] && k < 3 && (t[p] & 0xC0u) == 0x80u) { ++p; ++k; }
    }
    if (p >= e) return 0;
    int64_t nr = ak_row_lower_bound(off, r_lo, r_hi, p);       // next row-start event
    int64_t rs = (off[nr] == p) ? p : off[nr - 1];
    int64_t re = (off[nr] == p) ? p : off[nr];                  // fixed below when the row-start event fires
    AkCollapse st;
    st.c = 0xFFFFFFFFu;
    st.n = 0;
    st.pending = false;
    int64_t inert_until = -1;
    if (p != rs && p < total_end) {
        // mid-row start: which NFC segment are we in, and what did the row keep so far?
        int len;
        uint32_t cp = ak_decode(t, p, re, len);
        if (nfc && !AK_NFC_HEAD(ak_props(T, cp))) {
            int64_t h = ak_find_head(T, t, p, rs, re, limit, status);
            bool trouble;
            int64_t hend = ak_scan_segment(T, t, h, re, trouble, limit, status);
            if (trouble) p = hend;          // emitted by the span that owns the head
            else inert_until = hend;
        }
        if (p >= e) return 0;
        if (clean && p < re) {
            uint32_t k[3];
            int nk = ak_prev_kept(T, t, p, rs, re, flags, limit, k, status);
            if (nk > 0) {
                st.c = k[0];
                st.n = 1;
                if (k[0] != 0x0Au && nk > 1 && k[1] == k[0]) st.n = (nk > 2 && k[2] == k[0]) ? 3 : 2;
            }
        }
        if (p == re) {
            // we skipped a troubled segment that runs to the row end; its owner flushes the row
            nr = ak_row_lower_bound(off, nr, r_hi, p);
        }
    }
    for (;;) {
        if (p >= e) break;
        while (nr <= n_rows && off[nr] == p) {     // row-start events (several for empty rows)
            if (out_off) out_off[nr] = out_base + sink.cnt;
            ++nr;
            rs = p;
            st.c = 0xFFFFFFFFu;
            st.n = 0;
            st.pending = false;
            inert_until = -1;
        }
        if (p >= total_end) break;
        re = off[nr];
        int len;
        uint32_t cp = ak_decode(t, p, re, len);
        uint32_t w = ak_props(T, cp);
        bool slow = false;
        int64_t hend = p + len;
        if (nfc && p >= inert_until && (p == rs || AK_NFC_HEAD(w))) {
            // segment start: peek at the next code point; only scan when it is not itself a head
            bool trouble = AK_QC(w) != 0;
            if (hend < re) {
                int l2;
                uint32_t c2 = ak_decode(t, hend, re, l2);
                if (!AK_NFC_HEAD(ak_props(T, c2))) {
                    int64_t se = ak_scan_segment(T, t, p, re, trouble, limit, status);
                    if (trouble) hend = se; else inert_until = se;
                }
            }
            slow = trouble;
        }
        if (!slow) {
            uint32_t o[2];
            int m = ak_post_nfc(T, cp, w, flags, o);
            for (int i = 0; i < m; ++i) {
                if (clean) ak_collapse_feed(st, o[i], sink); else ak_sink_put(sink, o[i]);
            }
            p += len;
        } else {
            uint32_t buf[AK_MAXSEG];
            int n = ak_nfc_segment(T, t, p, hend, buf, status);
            for (int j = 0; j < n; ++j) {
                uint32_t o[2];
                int m = ak_post_nfc(T, buf[j], ak_props(T, buf[j]), flags, o);
                for (int i = 0; i < m; ++i) {
                    if (clean) ak_collapse_feed(st, o[i], sink); else ak_sink_put(sink, o[i]);
                }
            }
            p = hend;
        }
        if (p >= re) ak_collapse_close(st, 0xFFFFFFFFu, sink);      // the row ended with a code point of this span
    }
    // the span ended inside a row while holding back the 2nd element of a run: look ahead for the next kept one
    if (st.pending) ak_collapse_close(st, ak_next_kept(T, t, p, rs, re, flags, limit, status), sink);
    return sink.cnt;
}

// =================================================================================================
// grapheme clusters (reference segment.py:14,40-125 = regex \X, UAX #29) and script runs (segment.py:128-201)
// =================================================================================================
struct AkGState {
    uint8_t prev;     // GCB of the previous code point
    uint8_t conj;     // GB9c: 0 none, 1 Consonant [Extend|Linker]*, 2 ... with a Linker
    uint8_t pict;     // GB11: 0 none, 1 ExtPict Extend*, 2 ExtPict Extend* ZWJ
    uint8_t ri_odd;   // GB12/13: odd number of RI immediately before
    uint8_t prev_m;   // previous code point is a matra / halant (reference segment.py:20-37,84)
    uint8_t has_prev;
};

AK_HD bool ak_is_matra_or_halant(uint32_t cp) {
    return (cp >= 0x0900u && cp <= 0x0902u) || (cp >= 0x093Eu && cp <= 0x094Du) || (cp >= 0x0951u && cp <= 0x0954u);
}

AK_HD bool ak_g_break(const AkGState& st, uint32_t wb) {
    uint32_t ga = st.prev, gb = AK_GCB(wb);
    if (ga == GCB_CR && gb == GCB_LF) return false;                                              // GB3
    if (ga == GCB_CONTROL || ga == GCB_CR || ga == GCB_LF) return true;                          // GB4
    if (gb == GCB_CONTROL || gb == GCB_CR || gb == GCB_LF) return true;                          // GB5
    if (ga == GCB_L && (gb == GCB_L || gb == GCB_V || gb == GCB_LV || gb == GCB_LVT)) return false;   // GB6
    if ((ga == GCB_LV || ga == GCB_V) && (gb == GCB_V || gb == GCB_T)) return false;             // GB7
    if ((ga == GCB_LVT || ga == GCB_T) && gb == GCB_T) return false;                             // GB8
    if (gb == GCB_EXTEND || gb == GCB_ZWJ) return false;                                         // GB9
    if (gb == GCB_SPACINGMARK) return false;                                                     // GB9a
    if (ga == GCB_PREPEND) return false;                                                         // GB9b
    if (st.conj == 2 && AK_INCB(wb) == INCB_CONSONANT) return false;                             // GB9c
    if (st.pict == 2 && AK_EXTPICT(wb)) return false;                                            // GB11
    if (ga == GCB_RI && gb == GCB_RI && st.ri_odd) return false;                                 // GB12/13
    return true;                                                                                 // GB999
}

AK_HD void ak_g_advance(AkGState& st, uint32_t cp, uint32_t wb) {
    uint32_t gb = AK_GCB(wb), ib = AK_INCB(wb);
    st.ri_odd = (gb == GCB_RI) ? (st.ri_odd ^ 1) : 0;
    if (ib == INCB_CONSONANT) st.conj = 1;
    else if (ib == INCB_LINKER) st.conj = st.conj ? 2 : 0;
    else if (ib != INCB_EXTEND) st.conj = 0;
    if (AK_EXTPICT(wb)) st.pict = 1;
    else if (gb == GCB_EXTEND) st.pict = (st.pict == 1) ? 1 : 0;
    else if (gb == GCB_ZWJ) st.pict = (st.pict == 1) ? 2 : 0;
    else st.pict = 0;
    st.prev = (uint8_t)gb;
    st.prev_m = ak_is_matra_or_halant(cp) ? 1 : 0;
    st.has_prev = 1;
}

// the state after such a code point does not depend on what came before it
AK_HD bool ak_g_sync(uint32_t w) {
    uint32_t g = AK_GCB(w), ib = AK_INCB(w);
    return g != GCB_EXTEND && g != GCB_ZWJ && g != GCB_RI && ib != INCB_LINKER && ib != INCB_EXTEND;
}

struct AkSegOut {
    int32_t* cluster_ends;     // row-relative byte offset of every cluster (or cluster part) end
    int64_t* cluster_splits;   // [n_rows + 1]
    int32_t* run_ends;         // row-relative byte offset of every script-run end
    uint8_t* run_tags;         // TAG_* of the run, TAG_NONE for an all punct/digit row
    int64_t* run_splits;       // [n_rows + 1]
    int64_t cbase, rbase;      // global index of this span's first cluster / run
    int64_t ccap, rcap;        // capacities (writes beyond are dropped; counts stay exact)
    // AKSHAR_SEG_MASK: instead of the arrays above, the events of a (32-byte) lane as bits relative to lane_base:
    // lane_masks[0] cluster ends, [1] run ends, [2] / [3] the two tag planes (see ak_seg_kernels.cuh)
    uint32_t* lane_masks = nullptr;
    int64_t lane_base = 0;
};

AK_HD void ak_seg_put_cluster(const AkSegOut& o, bool write, int64_t cc, int64_t p, int64_t rs) {
    if (!write) return;
    if (o.lane_masks) { o.lane_masks[0] |= 1u << (int)(p - o.lane_base); return; }
    if (o.cbase + cc < o.ccap) o.cluster_ends[o.cbase + cc] = (int32_t)(p - rs);
}
AK_HD void ak_seg_put_run(const AkSegOut& o, bool write, int64_t rc, int64_t p, int64_t rs, uint32_t cur) {
    if (!write) return;
    if (o.lane_masks) {
        const uint32_t b = 1u << (int)(p - o.lane_base);
        o.lane_masks[1] |= b;
        if (cur == (uint32_t)TAG_ROMAN || cur == (uint32_t)TAG_NONE) o.lane_masks[2] |= b;
        if (cur == (uint32_t)TAG_OTHER || cur == (uint32_t)TAG_NONE) o.lane_masks[3] |= b;
        return;
    }
    if (o.rbase + rc < o.rcap) {
        o.run_ends[o.rbase + rc] = (int32_t)(p - rs);
        o.run_tags[o.rbase + rc] = (uint8_t)cur;
    }
}

AK_HD_NOINLINE void ak_seg_span(const AkTables& T, const uint8_t* t, const int64_t* off, int64_t n_rows, int64_t r_lo,
                                int64_t r_hi, int64_t s, int64_t e, uint32_t flags, int64_t limit, bool write,
                                const AkSegOut& o, int64_t& n_clusters, int64_t& n_runs, uint32_t& status) {
    const int64_t total_end = off[n_rows];
    const bool want_c = (flags & AK_SEG_CLUSTERS) != 0, want_r = (flags & AK_SEG_RUNS) != 0;
    const bool matras = (flags & AK_SEG_MATRAS) != 0;
    int64_t cc = 0, rc = 0;
    n_clusters = 0;
    n_runs = 0;
    int64_t p = s;
    if (p < total_end && p > off[0]) {
        // (only inside a row: a row that BEGINS with continuation bytes -- bytes that are not UTF-8 -- keeps them, so that
        // its row-start event is still met; nothing is skipped across the next row start either)
        const int64_t nr0 = ak_row_lower_bound(off, r_lo, r_hi, p);
        int k = 0;
        while (off[nr0] != s && p < e && p < total_end && p < off[nr0] && k < 3 && (t[p] & 0xC0u) == 0x80u) { ++p; ++k; }
    }
    if (p >= e) return;
    int64_t nr = ak_row_lower_bound(off, r_lo, r_hi, p);
    int64_t rs = (off[nr] == p) ? p : off[nr - 1];
    int64_t re = (off[nr] == p) ? p : off[nr];
    AkGState g;
    g.prev = 0; g.conj = 0; g.pict = 0; g.ri_odd = 0; g.prev_m = 0; g.has_prev = 0;
    uint32_t cur = TAG_NONE;
    // A row is closed (its last cluster end, its last run) by the span that owns the byte position where the row
    // ENDS, i.e. the start of the next row (or the end of the text) -- the same rule as the fast lane (ak_seg_fast.cuh).
    // A span that begins exactly there needs the closing row's run label: look back through that row.
    int64_t back_lo = rs;                  // start of the row whose state must be reconstructed
    bool need_cur = p != rs && p < total_end;
    if (p == rs && nr > 0 && off[nr - 1] < p) {
        rs = off[nr - 1];                  // the row that ends at p; the row-start event below closes it
        back_lo = rs;
        need_cur = true;
    } else if (p != rs && p < total_end) {
        if (want_c) {
            // back up to the nearest code point after which the rule state is history-free, then replay
            int64_t q = p;
            for (;;) {
                q = ak_prev_start(t, q, rs);
                int len;
                uint32_t cp = ak_decode(t, q, re, len);
                if (ak_g_sync(ak_props(T, cp))) break;
                if (q == rs) break;
                if (limit > 0 && p - q > limit) { status |= AK_ST_PATHOLOGICAL; break; }
            }
            while (q < p) {
                int len;
                uint32_t cp = ak_decode(t, q, re, len);
                ak_g_advance(g, cp, ak_props(T, cp));
                q += len;
            }
        }
    }
    if (want_r && need_cur) {
        int64_t q = p;
        while (q > back_lo) {
            q = ak_prev_start(t, q, back_lo);
            int len;
            uint32_t cp = ak_decode(t, q, p, len);
            uint32_t tg = AK_TAG(ak_props(T, cp));
            if (tg != TAG_DIGIT && tg != TAG_PUNCT) { cur = tg; break; }
            if (limit > 0 && p - q > limit) { status |= AK_ST_PATHOLOGICAL; break; }
        }
    }
    for (;;) {
        if (p >= e) break;
        while (nr <= n_rows && off[nr] == p) {
            if (nr > 0 && p > rs) {        // the row in progress is not empty: it ends here
                if (want_c) {
                    ak_seg_put_cluster(o, write, cc, p, rs);
                    ++cc;
                }
                if (want_r) {
                    ak_seg_put_run(o, write, rc, p, rs, cur);
                    ++rc;
                }
            }
            if (write && !o.lane_masks) {
                if (want_c) o.cluster_splits[nr] = o.cbase + cc;
                if (want_r) o.run_splits[nr] = o.rbase + rc;
            }
            ++nr;
            rs = p;
            g.prev = 0; g.conj = 0; g.pict = 0; g.ri_odd = 0; g.prev_m = 0; g.has_prev = 0;
            cur = TAG_NONE;
        }
        if (p >= total_end) break;
        re = off[nr];
        int len;
        uint32_t cp = ak_decode(t, p, re, len);
        uint32_t w = ak_props(T, cp);
        if (want_c) {
            if (g.has_prev) {
                bool brk = ak_g_break(g, w);
                if (matras && (g.prev_m || ak_is_matra_or_halant(cp))) brk = true;
                if (brk) {
                    ak_seg_put_cluster(o, write, cc, p, rs);
                    ++cc;
                }
            }
            ak_g_advance(g, cp, w);
        }
        if (want_r) {
            uint32_t tg = AK_TAG(w);
            if (tg != TAG_DIGIT && tg != TAG_PUNCT) {
                if (cur != TAG_NONE && tg != cur) {
                    ak_seg_put_run(o, write, rc, p, rs, cur);
                    ++rc;
                }
                cur = tg;
            }
        }
        p += len;
    }
    n_clusters = cc;
    n_runs = rc;
}

// =================================================================================================
// roman_phonetic_signature (reference normalize.py:59-89): lower() [all scripts, Final_Sigma] -> collapse runs
// >= 3 -> ee$ -> i, oo$ -> u -> aa kh gh ch th ph bh dh, on a code-point scratch `a` (one slot per input byte)
// =================================================================================================
AK_HD bool ak_final_sigma(const AkTables& T, const uint8_t* t, int64_t p, int len, int64_t rs, int64_t re) {
    // CPython handle_capital_sigma: \p{cased}\p{case-ignorable}* SIGMA !(\p{case-ignorable}* \p{cased})
    int64_t q = p;
    bool before = false;
    while (q > rs) {
        q = ak_prev_start(t, q, rs);
        int l;
        uint32_t w = ak_props(T, ak_decode(t, q, re, l));
        if ((w >> 26) & 1u) continue;
        before = ((w >> 27) & 1u) != 0;
        break;
    }
    if (!before) return false;
    q = p + len;
    while (q < re) {
        int l;
        uint32_t w = ak_props(T, ak_decode(t, q, re, l));
        q += l;
        if ((w >> 26) & 1u) continue;
        return ((w >> 27) & 1u) == 0;
    }
    return true;
}

AK_HD int ak_replace2(uint32_t* a, int n, uint32_t x, uint32_t y, uint32_t r) {
    int m = 0;
    for (int i = 0; i < n;) {
        if (i + 1 < n && a[i] == x && a[i + 1] == y) { a[m++] = r; i += 2; }
        else a[m++] = a[i++];
    }
    return m;
}

AK_HD_NOINLINE int ak_signature_row(const AkTables& T, const uint8_t* t, int64_t rs, int64_t re, uint32_t* a) {
    int n = 0;
    int64_t p = rs;
    // lower() then run collapse, streaming
    uint32_t last = 0xFFFFFFFFu;
    int run = 0;
#define AK_SIG_FEED(c_)                                   \
    do {                                                  \
        uint32_t c = (c_);                                \
        if (c == last) {                                  \
            if (c == 0x0Au) a[n++] = c;                   \
            else if (run < 3) ++run;                      \
        } else {                                          \
            if (run == 2) a[n++] = last;                  \
            a[n++] = c;                                   \
            last = c;                                     \
            run = 1;                                      \
        }                                                 \
    } while (0)
    while (p < re) {
        int len;
        uint32_t cp = ak_decode(t, p, re, len);
        uint32_t w = ak_props(T, cp);
        if (cp == 0x3A3u) AK_SIG_FEED(ak_final_sigma(T, t, p, len, rs, re) ? 0x3C2u : 0x3C3u);
        else if (AK_FULL_LOWER(w) || (cp >= 'A' && cp <= 'Z')) {
            AK_SIG_FEED(ak_full_lower(T, cp));
            if (cp == 0x130u) AK_SIG_FEED(0x307u);
        } else AK_SIG_FEED(cp);
        p += len;
    }
    if (run == 2) a[n++] = last;
#undef AK_SIG_FEED
    // ee$ -> i, oo$ -> u   (`$` also matches before one trailing newline)
    for (int pass = 0; pass < 2; ++pass) {
        const uint32_t v = pass == 0 ? 'e' : 'o', r = pass == 0 ? 'i' : 'u';
        if (n >= 2 && a[n - 1] == v && a[n - 2] == v) { a[n - 2] = r; --n; }
        else if (n >= 3 && a[n - 1] == 0x0Au && a[n - 2] == v && a[n - 3] == v) { a[n - 3] = r; a[n - 2] = 0x0Au; --n; }
    }
    n = ak_replace2(a, n, 'a', 'a', 'a');
    n = ak_replace2(a, n, 'k', 'h', 'k');
    n = ak_replace2(a, n, 'g', 'h', 'g');
    n = ak_replace2(a, n, 'c', 'h', 'c');
    n = ak_replace2(a, n, 't', 'h', 't');
    n = ak_replace2(a, n, 'p', 'h', 'p');
    n = ak_replace2(a, n, 'b', 'h', 'b');
    n = ak_replace2(a, n, 'd', 'h', 'd');
    return n;
}

