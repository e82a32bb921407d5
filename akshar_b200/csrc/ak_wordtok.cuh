// Word tokenizers (reference segment.py:239-401: word_tokenize_hindi / word_tokenize_sanskrit / word_tokenize) as parallel
// bit streams, 32 text bytes per lane.
//
// The reference walks the (already normalized) text one character at a time: `isspace` ends a word, danda / double danda
// (U+0964 U+0965) end a word and are tokens of their own, `.,!?;:()[]{}"'` end a word and are dropped, anything else
// joins the current word (segment.py:270-297; the Sanskrit routine, :335-362, is the same loop).  `word_tokenize` with
// language 'auto' takes that route when the raw text holds a code point of U+0900-097F and `text.split()` otherwise
// (segment.py:384-393): AKW_MODE_SPLIT is that second rule -- only `isspace` separates.
//
// Here a lane classifies its 32 bytes from the basis planes (ak_bits.cuh), the class of the previous code point rides the
// carry to the next lead, and the tokens fall out as two masks:
//   T  token starts: a word character whose predecessor in the row is not one, or a danda
//   E  token ends (exclusive): the lead / row start / end of text that follows a token's last code point
// The k-th set bit of T and the k-th set bit of E (in text order) delimit token k.  All helpers are AK_HD:
// tests/csrc/host_harness.cpp runs the identical arithmetic lane by lane on the CPU.
#pragma once
#include "ak_bits.cuh"
#include "ak_text_core.cuh"

#define AKW_MODE_HINDI 0
#define AKW_MODE_SPLIT 1

// truth table over three planes (index = 4 a + 2 b + c) and over a nibble (index = 8 p3 + 4 p2 + 2 p1 + p0)
template <uint32_t TT>
AK_HD uint32_t akb_lut3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if ((TT >> k) & 1u) r |= ((k & 4) ? a : ~a) & ((k & 2) ? b : ~b) & ((k & 1) ? c : ~c);
    return r;
}
template <uint32_t SET>
AK_HD uint32_t akb_nibble(uint32_t p3, uint32_t p2, uint32_t p1, uint32_t p0) {
    return (~p3 & akb_lut3<(SET & 0xFFu)>(p2, p1, p0)) | (p3 & akb_lut3<((SET >> 8) & 0xFFu)>(p2, p1, p0));
}

// str.isspace() of a code point beyond ASCII (Unicode 15: bidirectional class WS / B / S or category Zs)
AK_HD bool akw_isspace_wide(uint32_t cp) {
    return cp == 0x85u || cp == 0xA0u || cp == 0x1680u || (cp >= 0x2000u && cp <= 0x200Au) || cp == 0x2028u || cp == 0x2029u ||
           cp == 0x202Fu || cp == 0x205Fu || cp == 0x3000u;
}

struct AkWtLane {
    uint32_t own, rows, endbit;        // in: bytes of the text in this lane, row starts, the position text_end (if here)
    uint32_t cont, lead, SP, PU, E0b, A5b, A45b, A4or5b, hl;
    uint32_t DA, DEV, W;
    uint32_t dn;                       // out of phase 1: what the previous lane needs of my first two bytes
    uint32_t up;                       // out of phase 2: bit 0 last lead is a word character, 1 is a danda, 2 there is a lead
    uint32_t T, E;
};

AK_HD void akwt_phase1(const uint32_t* x, AkWtLane& L) {
    uint32_t P[8];
    akb_planes(x, P);
    const uint32_t p0 = P[0], p1 = P[1], p2 = P[2], p3 = P[3], p4 = P[4], p5 = P[5], p6 = P[6], p7 = P[7];
    const uint32_t asc = ~p7;
    L.cont = p7 & ~p6 & L.own;
    L.lead = ~(p7 & ~p6) & L.own;
    L.hl = p7 & p6 & L.own;
    // ASCII isspace: 09-0D 1C-1F 20
    const uint32_t c0 = asc & ~p6 & ~p5;
    L.SP = ((c0 & ~p4 & akb_nibble<0x3E00u>(p3, p2, p1, p0)) | (c0 & p4 & p3 & p2) | (asc & ~p6 & p5 & ~(p4 | p3 | p2 | p1 | p0))) & L.own;
    // . , ! ? ; : ( ) [ ] { } " '   = 21 22 27 28 29 2C 2E | 3A 3B 3F | 5B 5D | 7B 7D
    const uint32_t n2 = akb_nibble<0x5386u>(p3, p2, p1, p0), n3 = akb_nibble<0x8C00u>(p3, p2, p1, p0),
                   n57 = akb_nibble<0x2800u>(p3, p2, p1, p0);
    L.PU = asc & ((~p6 & p5 & ((~p4 & n2) | (p4 & n3))) | (p6 & p4 & n57)) & L.own;
    L.E0b = L.hl & p5 & ~p4 & ~(p3 | p2 | p1 | p0);
    const uint32_t a45 = p7 & ~p6 & p5 & ~p4 & ~p3 & p2 & ~p1;          // A4 / A5
    L.A4or5b = a45;
    L.A5b = a45 & p0;
    L.dn = (a45 & 3u) | ((L.A5b & 1u) << 2);
}

// dnn: the next lane's `dn` (0 past the text)
AK_HD void akwt_phase2(AkWtLane& L, uint32_t dnn, int mode) {
    const uint32_t n1_a45 = akb_fsr(L.A4or5b, dnn, 1);                  // byte + 1 is A4 / A5
    const uint32_t n1_a5 = akb_fsr(L.A5b, dnn >> 2, 1);                 // byte + 1 is A5
    const uint32_t n2_a45 = akb_fsr(L.A4or5b, dnn, 2);                  // byte + 2 is A4 / A5
    L.DEV = L.E0b & n1_a45;                                             // a code point of U+0900-097F
    L.DA = L.E0b & n1_a5 & n2_a45;                                      // U+0964 U+0965
    uint32_t sep = L.SP;
    if (mode == AKW_MODE_HINDI) sep |= L.PU | L.DA;
    else L.DA = 0;
    L.W = L.lead & ~sep;
}

// leads of two- to four-byte code points outside the Devanagari block: `isspace` needs the code point (rare)
AK_HD void akwt_wide(const uint8_t* text, int64_t cs, int64_t te, AkWtLane& L) {
    for (uint32_t m = L.hl & ~L.DEV; m;) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        int len;
        const uint32_t cp = ak_decode(text, cs + i, te, len);
        if (akw_isspace_wide(cp)) {
            L.SP |= 1u << i;
            L.W &= ~(1u << i);
        }
    }
}

AK_HD void akwt_summary(AkWtLane& L) {
    uint32_t up = 0;
    if (L.lead) {
        const uint32_t ll = 0x80000000u >> akb_clz(L.lead);
        up = ((L.W & ll) ? 1u : 0u) | ((L.DA & ll) ? 2u : 0u) | 4u;
    }
    L.up = up;
}

// upp: the previous lane's summary.  Returns the number of tokens open at the lane's first byte (0 / 1).
AK_HD uint32_t akwt_phase3(AkWtLane& L, uint32_t upp) {
    const uint32_t C = L.cont;
    const uint32_t pw = akb_fwd(L.W, C, upp & 1u);                      // the previous code point is a word character
    const uint32_t pd = akb_fwd(L.DA, C, (upp >> 1) & 1u);              //                        ... is a danda
    const uint32_t valid = L.own | L.endbit;
    L.T = ((L.W & ~(pw & ~L.rows)) | L.DA) & L.own;
    L.E = ((pw & (~L.W | L.rows)) | pd) & valid;
    return (upp & 3u) ? 1u : 0u;
}

// What a lane writes: the split of every row that starts in it (first_row .. first_row + nrows - 1), begin / end of its
// tokens (offsets relative to the row start), the Devanagari flag of its rows.  t_at / e_at = tokens that start / end
// before the lane, row_before = the last row that starts before the lane (-1: none).
AK_HD void akwt_emit_lane(const AkWtLane& L, int64_t cs, const int64_t* off, int64_t n_rows, int64_t first_row, int nrows,
                          int64_t row_before, int64_t t_at, int64_t e_at, int32_t* begin, int32_t* end, int64_t cap,
                          int64_t* splits, uint8_t* row_flags, uint32_t& st) {
    for (int j = 0; j < nrows; ++j) {
        const int64_t r = first_row + j;
        const int bit = (int)(off[r] - cs);
        splits[r] = t_at + akb_popc(L.T & ((1u << bit) - 1u));
    }
    const uint32_t ev = L.T | L.E;
    if (ev) {
        const uint32_t rows = L.rows;
        int64_t cur = 0;                                                // start of the row the walk is in
        if (!rows || akb_ctz(rows) >= akb_ctz(ev)) cur = off[row_before < 0 ? 0 : row_before];
        int64_t kt = t_at, ke = e_at;
        for (uint32_t m = ev | rows; m;) {
            const int i = akb_ctz(m);
            const uint32_t b = 1u << i;
            m &= m - 1u;
            const int64_t p = cs + i;
            if (L.E & b) {                                              // a token that ends where a row starts is the row before's
                if (ke >= 0 && ke < cap) end[ke] = (int32_t)(p - cur);
                else st |= AK_ST_OVERFLOW;
                ++ke;
            }
            if (rows & b) cur = p;
            if (L.T & b) {
                if (kt < cap) begin[kt] = (int32_t)(p - cur);
                else st |= AK_ST_OVERFLOW;
                ++kt;
            }
        }
    }
    if (row_flags && L.DEV) {
        int64_t r = row_before;
        int lo = 0;
        for (int j = 0; j <= nrows; ++j) {
            const int hi = j < nrows ? (int)(off[first_row + j] - cs) : 32;
            const uint32_t seg = (hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
            if ((L.DEV & seg) && r >= 0 && r < n_rows) row_flags[r] = 1;
            if (j < nrows) { r = first_row + j; lo = hi; }
        }
    }
}
