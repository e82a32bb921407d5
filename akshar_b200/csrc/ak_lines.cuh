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
