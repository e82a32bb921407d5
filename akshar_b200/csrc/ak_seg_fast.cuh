// Fast lane of segment_akshars / detect_code_switches (reference segment.py:40-201): UAX #29 extended grapheme
// clusters (`regex` \X, incl. GB9c / GB11 / GB12-13), the matras=True split, and the script-run grouping.
//
// Same chunking as ak_fast.cuh (16 bytes per lane, 30 real + 2 halo lanes per warp).  Phase A runs both little state
// machines over the lane's code points from a RESET state and summarises the chunk; the decisions that depend on
// what came before the chunk are exactly those up to the first code point after which the grapheme state is
// history-free (ak_g_sync) -- phase B re-evaluates only them from the previous chunk's end state -- and the run label
// in effect before the chunk's first strong-script code point, which is patched in from the previous chunk.  A
// chunk whose context is not determined by its left neighbour alone (no sync point in the neighbour: a run of
// regional indicators / ZWJ / marks; 16 bytes of digits and punctuation) takes the slow lane = the walker
// ak_seg_span, which implements the same ownership (events belong to the chunk that holds their byte position).
#pragma once
#include "ak_fast.cuh"

#define AKS_EXACT_G 1u         // end_g does not depend on the state before the chunk
#define AKS_ROWSTART 2u
#define AKS_USES_IN 4u         // some emission of this chunk carries the run label that was in effect before the chunk
#define AKS_CUR_IN 7u          // run label code: "whatever it was before the chunk"
#define AKS_CUR_NONE 5u        // no strong-script code point in the row yet

struct AkSChunk {
    uint32_t w[5];
    uint32_t rows, own;
    uint32_t lead;
    uint32_t brk;              // bit i: boundary before the code point led at byte i (row starts excluded)
    uint32_t fix;              // leads whose brk bit depends on the incoming grapheme state
    uint32_t rchg;             // bit i: a script run ends before the code point led at byte i
    uint32_t tag_lo, tag_hi;   // 3 bits per byte position: run label in effect just before position i
    uint32_t end_cur;          // run label code at the chunk end
    uint32_t first_strong;     // (pos << 8) | tag of the first strong code point seen while the label was AKS_CUR_IN, else AKF_NONE
    AkGState end_g;
    uint32_t flags;
};

AK_HD uint32_t aks_byte(const AkSChunk& c, int i) { return (c.w[i >> 2] >> ((i & 3) * 8)) & 0xFFu; }
AK_HD uint32_t aks_tag_before(const AkSChunk& c, int i) {
    return i < 10 ? (c.tag_lo >> (3 * i)) & 7u : (c.tag_hi >> (3 * (i - 10))) & 7u;
}
// GB3 - GB9b decide most pairs from (GCB before, GCB after) alone: one 32-bit word per GCB-before, low half = "no
// break" pairs, high half = pairs that always break (GB4 / GB5).  Built once per kernel from ak_g_break itself.
AK_HD uint32_t aks_pair_row(uint32_t ga) {
    uint32_t row = 0;
    AkGState g;
    g.prev = (uint8_t)ga; g.conj = 0; g.pict = 0; g.ri_odd = 0; g.prev_m = 0; g.has_prev = 1;
    for (uint32_t gb = 0; gb < 14u; ++gb) {
        const bool ctl_a = ga == GCB_CONTROL || ga == GCB_CR || ga == GCB_LF;
        const bool ctl_b = gb == GCB_CONTROL || gb == GCB_CR || gb == GCB_LF;
        if (!ak_g_break(g, gb)) row |= 1u << gb;                                   // props word with only the GCB field set
        else if ((ctl_a || ctl_b) && !(ga == GCB_CR && gb == GCB_LF)) row |= 1u << (16 + gb);
    }
    return row;
}
AK_HD bool aks_g_break(const uint32_t* gtab, const AkGState& st, uint32_t wb) {
    const uint32_t row = gtab[st.prev], gb = AK_GCB(wb);
    if ((row >> gb) & 1u) return false;
    if ((row >> (16 + gb)) & 1u) return true;
    if (st.conj == 2 && AK_INCB(wb) == INCB_CONSONANT) return false;                             // GB9c
    if (st.pict == 2 && AK_EXTPICT(wb)) return false;                                            // GB11
    if (st.prev == GCB_RI && gb == GCB_RI && st.ri_odd) return false;                            // GB12/13
    return true;
}

AK_HD void aks_g_reset(AkGState& g) { g.prev = 0; g.conj = 0; g.pict = 0; g.ri_odd = 0; g.prev_m = 0; g.has_prev = 0; }

AK_HD uint32_t aks_decode_at(const AkSChunk& c, int i, uint32_t b) {
    const uint32_t b1 = aks_byte(c, i + 1) & 0x3Fu, b2 = aks_byte(c, i + 2) & 0x3Fu, b3 = aks_byte(c, i + 3) & 0x3Fu;
    if (b < 0x80u) return b;
    if (b < 0xE0u) return ((b & 0x1Fu) << 6) | b1;
    if (b < 0xF0u) return ((b & 0x0Fu) << 12) | (b1 << 6) | b2;
    return ((b & 0x07u) << 18) | (b1 << 12) | (b2 << 6) | b3;
}

// ---- phase A -------------------------------------------------------------------------------------------------
// (loops over the byte positions are ROLLED: these kernels are instruction-cache bound)
AK_HD void aks_phase_a(const AkTables& T, const uint32_t* lut, AkSChunk& c, bool matras) {
    const uint32_t* gtab = lut + 384;      // 16 pair-rule rows follow the property table
    uint32_t lead = 0, brk = 0, fix = 0, rchg = 0, tag_lo = 0, tag_hi = 0, flags = 0;
    uint32_t cur = AKS_CUR_IN, first_strong = AKF_NONE;
    AkGState g;
    aks_g_reset(g);
    bool synced = false;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const uint32_t wlo = k == 0 ? c.w[0] : k == 1 ? c.w[1] : k == 2 ? c.w[2] : c.w[3];
        const uint32_t whi = k == 0 ? c.w[1] : k == 1 ? c.w[2] : k == 2 ? c.w[3] : c.w[4];
        const unsigned long long win = ((unsigned long long)whi << 32) | wlo;
        const uint32_t rows4 = (c.rows >> (4 * k)) & 15u, own4 = (c.own >> (4 * k)) & 15u;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int i = 4 * k + j;
            const bool row_here = (rows4 >> j) & 1u;
            // label in effect after everything before byte i (a row start at i closes its row with it)
            if (i < 10) tag_lo |= cur << (3 * i); else tag_hi |= cur << (3 * (i - 10));
            if (row_here) {
                if (cur == AKS_CUR_IN) flags |= AKS_USES_IN;
                aks_g_reset(g);
                synced = true;
                cur = AKS_CUR_NONE;
                flags |= AKS_ROWSTART;
            }
            const uint32_t v = (uint32_t)(win >> (8 * j));
            const uint32_t b = v & 0xFFu;
            if (!((own4 >> j) & 1u) || (b & 0xC0u) == 0x80u) continue;
            const uint32_t b1 = (v >> 8) & 0x3Fu, b2 = (v >> 16) & 0x3Fu, b3 = (v >> 24) & 0x3Fu;
            uint32_t cp;
            if (b < 0x80u) cp = b;
            else if (b < 0xE0u) cp = ((b & 0x1Fu) << 6) | b1;
            else if (b < 0xF0u) cp = ((b & 0x0Fu) << 12) | (b1 << 6) | b2;
            else cp = ((b & 0x07u) << 18) | (b1 << 12) | (b2 << 6) | b3;
            const uint32_t w = akf_props(T, lut, cp);
            lead |= 1u << i;
            // grapheme clusters
            if (!synced) fix |= 1u << i;
            if (g.has_prev) {
                bool bk = aks_g_break(gtab, g, w);
                if (matras && (g.prev_m || ak_is_matra_or_halant(cp))) bk = true;
                if (bk) brk |= 1u << i;
            }
            ak_g_advance(g, cp, w);
            if (ak_g_sync(w)) synced = true;
            // script runs
            const uint32_t t = AK_TAG(w);
            if (t != TAG_DIGIT && t != TAG_PUNCT) {
                if (cur == AKS_CUR_IN) { first_strong = ((uint32_t)i << 8) | t; flags |= AKS_USES_IN; }
                else if (cur != AKS_CUR_NONE && cur != t) rchg |= 1u << i;
                cur = t;
            }
        }
    }
    if (synced) flags |= AKS_EXACT_G;
    c.lead = lead;
    c.brk = brk;
    c.fix = fix;
    c.rchg = rchg;
    c.tag_lo = tag_lo;
    c.tag_hi = tag_hi;
    c.end_cur = cur;
    c.first_strong = first_strong;
    c.end_g = g;
    c.flags = flags;
}

struct AkSNeighbor {
    AkGState g;          // end state of the previous chunk
    uint32_t flags;
    uint32_t end_cur;
};

// ---- phase B: patch the context-dependent decisions; false = slow lane ----------------------------------------
AK_HD bool aks_phase_b(const AkTables& T, const uint32_t* lut, AkSChunk& c, const AkSNeighbor& prev, bool matras,
                       bool want_c, bool want_r, uint32_t& in_cur) {
    in_cur = prev.end_cur;
    const uint32_t* gtab = lut + 384;
    if (want_c && c.fix) {
        if (!(prev.flags & AKS_EXACT_G)) return false;
        AkGState g = prev.g;
        uint32_t m = c.fix;
        uint32_t brk = c.brk & ~c.fix;
        while (m) {
#ifdef __CUDA_ARCH__
            const int i = __ffs(m) - 1;
#else
            const int i = __builtin_ctz(m);
#endif
            m &= m - 1u;
            // dynamic byte position: fetch the (up to) 4 bytes at i through a 64-bit window
            const int wi = i >> 2, sh = (i & 3) * 8;
            const uint32_t lo = wi == 0 ? c.w[0] : wi == 1 ? c.w[1] : wi == 2 ? c.w[2] : c.w[3];
            const uint32_t hi = wi == 0 ? c.w[1] : wi == 1 ? c.w[2] : wi == 2 ? c.w[3] : c.w[4];
            const uint32_t v = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
            const uint32_t b = v & 0xFFu, b1 = (v >> 8) & 0x3Fu, b2 = (v >> 16) & 0x3Fu, b3 = (v >> 24) & 0x3Fu;
            uint32_t cp;
            if (b < 0x80u) cp = b;
            else if (b < 0xE0u) cp = ((b & 0x1Fu) << 6) | b1;
            else if (b < 0xF0u) cp = ((b & 0x0Fu) << 12) | (b1 << 6) | b2;
            else cp = ((b & 0x07u) << 18) | (b1 << 12) | (b2 << 6) | b3;
            const uint32_t w = akf_props(T, lut, cp);
            if (g.has_prev) {
                bool bk = aks_g_break(gtab, g, w);
                if (matras && (g.prev_m || ak_is_matra_or_halant(cp))) bk = true;
                if (bk) brk |= 1u << i;
            }
            ak_g_advance(g, cp, w);
        }
        c.brk = brk;
    }
    if (want_r) {
        if ((c.flags & AKS_USES_IN) && in_cur == AKS_CUR_IN) return false;
        if (c.first_strong != AKF_NONE && in_cur != AKS_CUR_NONE && in_cur != (c.first_strong & 0xFFu))
            c.rchg |= 1u << (c.first_strong >> 8);
    }
    return true;
}

// ---- emission -------------------------------------------------------------------------------------------------
struct AkSegSink {
    int32_t* cbuf;      // staged cluster ends (index i * stride)
    int32_t* rbuf;      // staged run ends
    uint8_t* tbuf;      // staged run tags
    int cap, stride;
    int cc, rc;
    bool direct;
    int32_t* gc;        // direct mode: global streams, already offset to this lane's base
    int32_t* gr;
    uint8_t* gt;
    int64_t gccap, grcap;    // elements this lane may still write in direct mode
};
AK_HD void aks_put_c(AkSegSink& s, int32_t v) {
    if (s.direct) { if (s.cc < s.gccap) s.gc[s.cc] = v; }
    else if (s.cc < s.cap) s.cbuf[(int64_t)s.cc * s.stride] = v;
    ++s.cc;
}
AK_HD void aks_put_r(AkSegSink& s, int32_t v, uint32_t tag) {
    if (s.direct) { if (s.rc < s.grcap) { s.gr[s.rc] = v; s.gt[s.rc] = (uint8_t)tag; } }
    else if (s.rc < s.cap) { s.rbuf[(int64_t)s.rc * s.stride] = v; s.tbuf[(int64_t)s.rc * s.stride] = (uint8_t)tag; }
    ++s.rc;
}

AK_HD uint32_t aks_label(uint32_t code, uint32_t in_cur) {
    if (code == AKS_CUR_IN) code = in_cur;
    return code == AKS_CUR_NONE ? (uint32_t)TAG_NONE : code;
}

// everything the lane emits, in stream order.  nr = index of the first row that starts at or after the lane's first
// position (so the row in progress is nr - 1); splits receive lane-relative indices for rows [row_first, row_last)
AK_HD_NOINLINE void aks_lane_emit(const AkSChunk& c, uint32_t in_cur, int64_t cs, const int64_t* off, int64_t n_rows, int64_t nr,
                                  bool want_c, bool want_r, AkSegSink& sink, int64_t* csplits, int64_t* rsplits,
                                  int64_t& row_first, int64_t& row_last) {
    uint32_t ev = c.rows & 0xFFFFu;
    if (want_c) ev |= c.brk;
    if (want_r) ev |= c.rchg;
    row_first = row_last = nr;
    int64_t rs = nr > 0 ? off[nr - 1] : off[0];
    while (ev) {
#ifdef __CUDA_ARCH__
        const int i = __ffs(ev) - 1;
#else
        const int i = __builtin_ctz(ev);
#endif
        ev &= ev - 1u;
        const int64_t p = cs + i;
        if ((c.rows >> i) & 1u) {
            if (nr > 0 && p > rs) {      // the row in progress is not empty: it ends here
                if (want_c) aks_put_c(sink, (int32_t)(p - rs));
                if (want_r) aks_put_r(sink, (int32_t)(p - rs), aks_label(aks_tag_before(c, i), in_cur));
            }
            while (nr <= n_rows && off[nr] == p) {
                if (csplits) csplits[nr] = sink.cc;
                if (rsplits) rsplits[nr] = sink.rc;
                ++nr;
            }
            row_last = nr;
            rs = p;
        } else {
            if (want_c && ((c.brk >> i) & 1u)) aks_put_c(sink, (int32_t)(p - rs));
            if (want_r && ((c.rchg >> i) & 1u)) aks_put_r(sink, (int32_t)(p - rs), aks_label(aks_tag_before(c, i), in_cur));
        }
    }
}
