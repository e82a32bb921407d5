// Front end of the BPE encoder (HF `Whitespace` pre-tokenizer `\w+|[^\w\s]+`, reference scripts/train_bpe.py:68-98 and
// tokenizer.py:193) as parallel bit streams, 32 text bytes per lane.
//
// From the basis planes (ak_bits.cuh): the pre-tokenizer class of every code point of the closed alphabet -- ASCII and
// U+0900-097F -- as three lead-byte masks (word / space / other), word boundaries where the class changes (the previous
// code point's class rides the carry to the next lead) or a row starts, word starts = boundaries that are not space,
// and the NFC trouble bits of ak_norm3.cuh (the encoder takes its text as is and only has to notice when NFC would
// change it).  Code points outside the closed alphabet are classified one by one from the property table (emoji,
// accents ... with clean_hinglish=False); `<` and the code points HF's NFKC treats differently raise AK_ST_ALPHABET.
#pragma once
#include "ak_bits.cuh"
#include "ak_text_core.cuh"
#include "ak_subword.cuh"

struct AkB3Lane {
    uint32_t own, rows;
    uint32_t cont, lead;
    uint32_t E0b, hl, A4b, A5b, P5b, NKb, NIb, R2b, R3b, QNb;      // byte roles
    uint32_t CW, CS;                 // class word / space at leads (other = lead & ~CW & ~CS)
    uint32_t NK, VIAC, AC, QN, INERT, FOR, LT;
    uint32_t bnd, wstart, trb;
    uint32_t UNS;                    // code points HF's NFKC treats differently from NFC: their rows take the exact row path
    uint32_t dn1, up2;
};

// up2: bits 0-1 class of the last lead (1 word, 2 space, 0 other), bit 2 lane has a lead, bit 3 INERT at the last lead,
// bit 4 accent / foreign mark at the last lead
AK_HD void akb3_phase1(const uint32_t* x, AkB3Lane& L) {
    uint32_t P[8];
    akb_planes(x, P);
    const uint32_t p0 = P[0], p1 = P[1], p2 = P[2], p3 = P[3], p4 = P[4], p5 = P[5], p6 = P[6], p7 = P[7];
    const uint32_t own = L.own;
    const uint32_t asc = ~p7 & own;
    const uint32_t cont = p7 & ~p6;
    L.cont = cont;
    L.lead = ~cont & own;
    L.hl = p7 & p6 & own;
    L.E0b = L.hl & p5 & ~p4 & ~(p3 | p2 | p1 | p0);
    const uint32_t a45 = cont & p5 & ~p4 & ~p3 & p2 & ~p1;
    L.A4b = a45 & ~p0;
    L.A5b = a45 & p0;
    // U+0964 0965 0970 (third bytes A4 A5 B0 after A5) are the block's only non-word code points
    L.P5b = cont & p5 & ~p3 & ~p1 & ((~p4 & p2) | (p4 & ~p2 & ~p0));
    const uint32_t c01 = cont & ~p5;
    L.NKb = cont & p5 & p4 & p3 & p2 & ~p1 & ~p0;
    L.R2b = c01 & ~p4 & p3 & p2 & ~p1 & p0;
    L.R3b = c01 & p4 & ~p3 & ((~p2 & (p1 | p0)) | (p2 & ~p1 & ~p0));
    L.QNb = c01 & p4 & p3;
    L.NIb = cont & p5 & ((~p4 & p3 & ~p2 & ~p1) | (p4 & ~p3 & ((~p2 & ~(p1 ^ p0)) | (~p2 & ~p1) | (p2 & ~p1 & ~p0))));
    // ASCII: word = 0-9 A-Z _ a-z, space = 09-0D 20
    const uint32_t low_nz = p4 | p3 | p2 | p1 | p0;
    const uint32_t low_gt26 = p4 & p3 & (p2 | (p1 & p0));
    const uint32_t letters = asc & p6 & low_nz & ~low_gt26;
    const uint32_t digits = asc & ~p6 & p5 & p4 & (~p3 | (~p2 & ~p1));
    const uint32_t under = asc & p6 & ~p5 & p4 & p3 & p2 & p1 & p0;
    L.CW = letters | digits | under;
    const uint32_t row0 = asc & ~p6 & ~p5 & ~p4;
    L.CS = (row0 & p3 & ((~p2 & (p1 | p0)) | (p2 & ~p1))) | (asc & ~p6 & p5 & ~(p4 | p3 | p2 | p1 | p0));
    L.LT = asc & ~p6 & p5 & p4 & p3 & p2 & ~p1 & ~p0;                 // '<'
    L.dn1 = (L.A4b & 1u) | ((L.A5b & 1u) << 1) | ((L.P5b & 3u) << 2) | ((L.NKb & 3u) << 4) | ((L.NIb & 3u) << 6) |
            ((L.R2b & 3u) << 8) | ((L.R3b & 3u) << 10) | ((L.QNb & 3u) << 12);
}

AK_HD void akb3_phase2(AkB3Lane& L, uint32_t dn1n) {
    const uint32_t a4_1 = akb_fsr(L.A4b, dn1n, 1), a5_1 = akb_fsr(L.A5b, dn1n >> 1, 1);
    const uint32_t d4 = L.E0b & a4_1, d5 = L.E0b & a5_1;
    const uint32_t dev = d4 | d5;
    L.FOR = L.hl & ~dev;
    L.CW |= dev & ~(d5 & akb_fsr(L.P5b, dn1n >> 2, 2));
    L.NK = d4 & akb_fsr(L.NKb, dn1n >> 4, 2);
    const uint32_t r2 = d5 & akb_fsr(L.R2b, dn1n >> 8, 2), r3 = d5 & akb_fsr(L.R3b, dn1n >> 10, 2);
    L.VIAC = r2 | r3;
    L.AC = r3;
    L.QN = d5 & akb_fsr(L.QNb, dn1n >> 12, 2);
    L.INERT = dev & ~(L.NK | L.VIAC | L.QN) & ~(d4 & akb_fsr(L.NIb, dn1n >> 6, 2));
    L.UNS = 0u;
}

// code points outside the closed alphabet, one by one
AK_HD void akb3_foreign(const AkTables& Tb, const uint8_t* text, int64_t cs, int64_t te, AkB3Lane& L) {
    uint32_t xt = 0;
    for (uint32_t m = L.FOR; m;) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        int len;
        const uint32_t cp = ak_decode(text, cs + i, te, len);
        const uint32_t w = ak_props(Tb, cp);
        const uint32_t k = AK_HFCLASS(w);
        if (k == 1u) L.CW |= 1u << i;
        else if (k == 2u) L.CS |= 1u << i;
        if (!AK_BPE_SAFE(w)) L.UNS |= 1u << i;
        if (!AK_INERT_BASE(w)) xt |= 1u << i;          // anything NFC might care about: checked exactly (cold)
    }
    L.QN |= xt;
    L.AC |= xt;
}

AK_HD void akb3_summary(AkB3Lane& L) {
    uint32_t up = 0;
    if (L.lead) {
        const uint32_t ll = 0x80000000u >> akb_clz(L.lead);
        up = ((L.CW & ll) ? 1u : (L.CS & ll) ? 2u : 0u) | 4u | ((L.INERT & ll) ? 8u : 0u) | ((L.AC & ll) ? 16u : 0u);
    }
    L.up2 = up;
}

// up2p: the previous lane's summary; for a lane without a lead (outside the text) it reads "nothing before"
AK_HD void akb3_phase3(AkB3Lane& L, uint32_t up2p) {
    const uint32_t C = L.cont;
    const uint32_t lead = L.lead;
    const uint32_t other = lead & ~L.CW & ~L.CS;
    const uint32_t has = (up2p >> 2) & 1u, pk = up2p & 3u;
    const uint32_t pw = akb_fwd(L.CW, C, (has && pk == 1u) ? 1u : 0u);
    const uint32_t ps = akb_fwd(L.CS, C, (has && pk == 2u) ? 1u : 0u);
    const uint32_t po = akb_fwd(other, C, (has && pk == 0u) ? 1u : 0u);
    // same class as the previous code point of the row -> no boundary
    const uint32_t same = ((L.CW & pw) | (L.CS & ps) | (other & po)) & ~L.rows;
    L.bnd = (lead & ~same) | L.rows;
    L.wstart = L.bnd & lead & ~L.CS;
    const uint32_t p_inert = akb_fwd(L.INERT, C, (up2p >> 3) & 1u) & ~L.rows;
    const uint32_t p_ac = akb_fwd(L.AC, C, (up2p >> 4) & 1u) & ~L.rows;
    L.trb = (L.QN | (L.NK & ~p_inert) | (L.VIAC & p_ac)) & L.own;
}

// exact NFC check of the troubled code points (cold): does NFC change the text?
AK_HD_NOINLINE bool akb3_changes(const AkTables& T, const uint8_t* text, const int64_t* off, int64_t n_rows, int64_t r_lo,
                                          uint32_t trb, int64_t cs, uint32_t& status) {
    bool changed = false;
    int64_t checked_until = -1;
    while (trb) {
        const int i = akb_ctz(trb);
        trb &= trb - 1u;
        const int64_t p = cs + i;
        if (p < checked_until) continue;
        const int64_t r = ak_row_lower_bound(off, r_lo, n_rows, p + 1);
        const int64_t rs = off[r - 1], re = off[r];
        if (ak_segment_changes(T, text, p, rs, re, 4096, &checked_until, status)) changed = true;
    }
    return changed;
}

// end of a word of class k with no boundary before `from` (cold: words longer than 64 bytes)
AK_HD_NOINLINE int64_t akb3_scan_end(const AkTables& T, const uint8_t* t, int64_t wpos, int64_t from, uint32_t k,
                                              const int64_t* off, int64_t n_rows, int64_t r_lo, int64_t r_hi) {
    int64_t er = ak_row_lower_bound(off, r_lo, r_hi, wpos + 1);
    if (off[er] < wpos + 1) er = ak_row_lower_bound(off, r_hi, n_rows, wpos + 1);
    const int64_t re = off[er];
    int64_t q = from;
    if (q > re) q = re;
    while (q < re && (t[q] & 0xC0u) == 0x80u) ++q;
    while (q < re) {
        int len;
        const uint32_t cp = ak_decode(t, q, re, len);
        if (AK_HFCLASS(ak_props(T, cp)) != k) break;
        q += len;
    }
    return q;
}

