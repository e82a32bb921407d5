// Lines of a text file as rows, on the device (reference cli.py:165-190 = scripts/train_bpe.py:16-35 = train_spm.py:18-44):
//     for line in f.readlines(): line = line.strip(); if not line: continue
// `f` is a text-mode file with universal newlines: '\n', '\r\n' and a lone '\r' end a line; `str.strip()` removes the code
// points `str.isspace()` accepts from both ends.  So a row begins at the first code point of a line that is not white space
// and ends after the last one; lines without one are skipped.
//
// A byte stream with a two-state machine ("has this line shown a kept code point yet?"): a kept code point in state 0
// BEGINS a row, a terminator in state 1 ENDS the row at the end of the last kept code point.  Every thread owns the code
// points whose lead byte falls into its 32 bytes and summarizes them as a transducer (next state and rows begun, for
// either entry state, plus the last kept position); transducers compose, so a warp scan, one small resolve kernel over the
// per-tile summaries and a second pass over the text place every row without any ordering between CTAs.
#pragma once
#include "ak_unicode.cuh"
#include "ak_wordtok.cuh"

#define AKLN_SPAN 32                                  // bytes per thread
#define AKLN_TILE (32 * AKLN_SPAN)                    // bytes per warp tile

// summary of a stretch of text: s = exit state for entry state 0 / 1 (bits 0 / 1), cnt[x] = rows begun for entry state x,
// lastk = position after the last kept code point (-1: none)
struct AkLineFn {
    uint32_t s;
    int32_t cnt0, cnt1;
    int64_t lastk;
};
AK_HD AkLineFn akl_identity() {
    AkLineFn f;
    f.s = 2u; f.cnt0 = 0; f.cnt1 = 0; f.lastk = -1;
    return f;
}
// a then b
AK_HD AkLineFn akl_compose(const AkLineFn& a, const AkLineFn& b) {
    AkLineFn r;
    const uint32_t a0 = a.s & 1u, a1 = (a.s >> 1) & 1u;
    r.s = ((b.s >> a0) & 1u) | (((b.s >> a1) & 1u) << 1);
    r.cnt0 = a.cnt0 + (a0 ? b.cnt1 : b.cnt0);
    r.cnt1 = a.cnt1 + (a1 ? b.cnt1 : b.cnt0);
    r.lastk = b.lastk >= 0 ? b.lastk : a.lastk;
    return r;
}

// class of the code point whose lead byte is at p: 0 kept, 1 white space (str.isspace), 2 line terminator
AK_HD int akl_class(const uint8_t* t, int64_t p, int64_t te, int& len) {
    const uint32_t b0 = t[p];
    if (b0 < 0x80u) {
        len = 1;
        if (b0 == 0x0Au || b0 == 0x0Du) return 2;
        return ((b0 >= 0x09u && b0 <= 0x0Du) || (b0 >= 0x1Cu && b0 <= 0x20u)) ? 1 : 0;
    }
    const uint32_t cp = ak_decode(t, p, te, len);
    return akw_isspace_wide(cp) ? 1 : 0;
}

// the code points whose lead byte is in [s, e): summary (emit == false), or with the entry state known the rows themselves:
// `state` / `lastk` carry in and out, `rank` = rows begun before s; begin[k] / end[k] receive absolute byte positions.
AK_HD AkLineFn akl_span(const uint8_t* t, int64_t s, int64_t e, int64_t te, bool emit, uint32_t state, int64_t lastk, int64_t rank,
                        int64_t* begin, int64_t* end, int64_t cap, uint32_t& st) {
    uint32_t s0 = 0u, s1 = 1u;
    int32_t c0 = 0, c1 = 0;
    int64_t lk = -1;
    for (int64_t p = s; p < e && p <= te; ++p) {
        int cls, len = 1;
        if (p == te) cls = 2;                                   // the end of the file ends the last line
        else {
            if ((t[p] & 0xC0u) == 0x80u) continue;
            cls = akl_class(t, p, te, len);
        }
        if (!emit) {
            if (cls == 0) {
                if (!s0) { s0 = 1u; ++c0; }
                if (!s1) { s1 = 1u; ++c1; }
                lk = p + len;
            } else if (cls == 2) s0 = s1 = 0u;
        } else {
            if (cls == 0) {
                if (!state) {
                    state = 1u;
                    if (rank < cap) begin[rank] = p; else st |= AK_ST_OVERFLOW;
                    ++rank;
                }
                lastk = p + len;
            } else if (cls == 2) {
                if (state) {
                    if (rank - 1 < cap) end[rank - 1] = lastk;
                    state = 0u;
                }
            }
        }
    }
    AkLineFn f;
    f.s = s0 | (s1 << 1);
    f.cnt0 = c0;
    f.cnt1 = c1;
    f.lastk = lk;
    return f;
}

// ---- the same transducer of a 32-byte lane from bit masks (the kernels' path; akl_span above is the byte-by-byte statement
// of it that the CPU tests also run over other span sizes) --------------------------------------------------------------
struct AkLnLane {
    uint32_t own, endbit;            // in: bytes of the file in this lane, the position n (end of file) if it is here
    uint32_t lead, K, TERM, WIDE;    // lead bytes; kept code points and terminators (at their lead bytes); leads to decode
};

AK_HD void akln_phase1(const uint32_t* x, AkLnLane& L) {
    uint32_t P[8];
    akb_planes(x, P);
    const uint32_t p0 = P[0], p1 = P[1], p2 = P[2], p3 = P[3], p4 = P[4], p5 = P[5], p6 = P[6], p7 = P[7];
    const uint32_t asc = ~p7;
    L.lead = ~(p7 & ~p6) & L.own;
    const uint32_t c0 = asc & ~p6 & ~p5;
    // ASCII isspace: 09-0D 1C-1F 20; terminators 0A 0D
    const uint32_t sp = (c0 & ~p4 & akb_nibble<0x3E00u>(p3, p2, p1, p0)) | (c0 & p4 & p3 & p2) | (asc & ~p6 & p5 & ~(p4 | p3 | p2 | p1 | p0));
    L.TERM = (c0 & ~p4 & akb_nibble<0x2400u>(p3, p2, p1, p0) & L.own) | L.endbit;
    // leads of the code points str.isspace() accepts beyond ASCII: C2 (85 A0), E1 (9A 80), E2 (80 .. / 81 9F), E3 (80 80)
    L.WIDE = p7 & p6 & L.own & ((~p5 & ~p4 & ~p3 & ~p2 & p1 & ~p0) | (p5 & ~p4 & ~p3 & ~p2 & (p1 | p0)));
    L.K = L.lead & ~sp;
}
AK_HD void akln_wide(const uint8_t* t, int64_t cs, int64_t te, AkLnLane& L) {
    for (uint32_t m = L.WIDE; m;) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        int len;
        if (akw_isspace_wide(ak_decode(t, cs + i, te, len))) L.K &= ~(1u << i);
    }
}
// position after the code point whose lead byte is bit i of the lane
AK_HD int64_t akln_cp_end(const uint8_t* t, int64_t cs, int i, int64_t te) {
    const uint32_t b = t[cs + i];
    int64_t e = cs + i + (b < 0x80u ? 1 : b < 0xE0u ? 2 : b < 0xF0u ? 3 : 4);
    return e > te ? te : e;
}
AK_HD AkLineFn akln_summary(const AkLnLane& L, const uint8_t* t, int64_t cs, int64_t te) {
    AkLineFn f;
    const uint32_t ev = L.K | L.TERM;
    if (!ev) return akl_identity();
    const uint32_t first0 = L.K & akb_fwd(L.TERM, ~ev, 1u), first1 = L.K & akb_fwd(L.TERM, ~ev, 0u);
    const uint32_t top = 0x80000000u >> akb_clz(ev);
    const uint32_t out = (L.K & top) ? 1u : 0u;
    f.s = out | (out << 1);
    f.cnt0 = akb_popc(first0);
    f.cnt1 = akb_popc(first1);
    f.lastk = L.K ? akln_cp_end(t, cs, 31 - akb_clz(L.K), te) : -1;
    return f;
}
AK_HD void akln_emit(const AkLnLane& L, const uint8_t* t, int64_t cs, int64_t te, uint32_t state, int64_t lastk, int64_t rank,
                     int64_t* begin, int64_t* end, int64_t cap, uint32_t& st) {
    const uint32_t ev = L.K | L.TERM;
    if (!ev) return;
    const uint32_t first = L.K & akb_fwd(L.TERM, ~ev, state ? 0u : 1u);
    const uint32_t ends = L.TERM & akb_fwd(L.K, ~ev, state ? 1u : 0u);
    for (uint32_t m = first; m;) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        const int64_t k = rank + akb_popc(first & ((1u << i) - 1u));
        if (k < cap) begin[k] = cs + i; else st |= AK_ST_OVERFLOW;
    }
    for (uint32_t m = ends; m;) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        const uint32_t below = L.K & ((1u << i) - 1u);
        const int64_t e = below ? akln_cp_end(t, cs, 31 - akb_clz(below), te) : lastk;
        const int64_t k = rank + akb_popc(first & ((1u << i) - 1u)) - 1;          // the row that is open at this terminator
        if (k >= 0 && k < cap) end[k] = e;
    }
}
