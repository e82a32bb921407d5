// segment_akshars / detect_code_switches (reference segment.py:40-201) as parallel bit streams, 32 text bytes per
// lane -- the hot kernel's fast lane for the alphabet normalize_text leaves behind: ASCII and U+0900-097F.
//
// From the basis planes (ak_bits.cuh) the lane derives, as masks over its byte positions,
//   grapheme classes   CR, LF, Control (ASCII); Extend / SpacingMark / Consonant / Linker (U+0900-097F, Unicode 17
//                      as `regex` \X sees it -- tests/test_span_walkers.py checks every byte role against the table)
//   brk                a cluster boundary before the code point led here: every lead except
//                      GB3 CR x LF, GB9 / GB9a x (Extend | SpacingMark) unless a control precedes (GB4),
//                      GB9c consonant (Extend | Linker)* Linker (Extend | Linker)* x consonant -- the two "conjunct
//                      state" sets are MatchStar runs: one integer add each;
//                      matras=True adds a boundary before and after every matra / halant (segment.py:80-125)
//   rchg               a script run ends before the strong (not digit / punct) code point led here: its label differs
//                      from the previous strong one's, which travels to it over the weak ones by carry propagation
//   PD / PR / PO       label of the previous strong code point, at every event position (run tags)
// Code points outside ASCII / U+0900-097F (Bengali, NBSP, emoji when the input is not normalized ...) are classified one
// by one from the property table and join the same masks: GB11 (ExtPict Extend* ZWJ x ExtPict) is one more MatchStar,
// GB12/13 a parity over runs of regional indicators.  Only the classes the masks do not model (Prepend, Hangul syllable
// types) make the lane slow = the walker ak_seg_span on its 32 bytes, and tell its successor not to trust its end state.
// Cross-lane context: one packed word down (look-ahead byte roles), one up (end-of-lane summaries).
#pragma once
#include "ak_bits.cuh"
#include "ak_text_core.cuh"

struct AkS3Lane {
    uint32_t own, rows;
    uint32_t cont, lead;
    // byte roles (phase 1)
    uint32_t E0b, A4b, A5b, hl;
    uint32_t X4b, S4b, C4b, X5b, S5b, C5b, LKb, M4b, M5b;     // third-byte roles after A4 / A5
    uint32_t CR, LF, CTL, ROM, WEAK;                            // ASCII leads
    // code-point masks (phase 2)
    uint32_t DEV, FOR, X, SM, CONS, LK, MAT, strong_d, strong_r, strong_o;
    // X = GCB Extend or ZWJ (GB9); XI = InCB Extend / Linker (transparent for GB9c); XE = GCB Extend only (transparent for
    // GB11); ZW / EP / RI = ZWJ, Extended_Pictographic, regional indicators; UNS = foreign code points of a class the masks
    // do not model (Prepend, Hangul syllable types): the lane then takes the walker
    uint32_t XI, XE, ZW, EP, RI, UNS;
    // results
    uint32_t brk, rchg, PD, PR, PO;
    uint32_t dn1, up2;
};

// up2 layout
#define AKS3_LAST_CTL 1u        // last lead is CR / LF / Control
#define AKS3_LAST_CR 2u
#define AKS3_LAST_MAT 4u
#define AKS3_S1 8u              // conjunct state after the last code point: >= 1 / == 2
#define AKS3_S2 16u
#define AKS3_G_OPEN 32u         // no lead that fixes the grapheme state (all Extend / Linker, no barrier), or a foreign one after it
#define AKS3_LAB_SHIFT 6        // 3 bits: 0 none (barrier, no strong after it), 1 dev, 2 roman, 3 other, 7 open (no strong, no barrier / foreign)
#define AKS3_LAST_UNS 512u      // last lead is of an unsupported class
#define AKS3_P1 1024u           // GB11 state after the last code point: ExtPict Extend* / ... ZWJ
#define AKS3_P2 2048u
#define AKS3_LAST_RI 4096u      // last lead is a regional indicator / an odd number of them ends the lane
#define AKS3_RI_ODD 8192u

// ---- phase 1: planes and byte roles ----------------------------------------------------------------------------
AK_HD void aks3_phase1(const uint32_t* x, AkS3Lane& L) {
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
    // third bytes: v = low six bits
    const uint32_t lo3z = ~(p2 | p1 | p0);
    // after A4 (U+0900 + v): Extend 00-02 3A 3C, SpacingMark 03 3B 3E 3F, consonant 15-39, matra 00-02 3E 3F
    const uint32_t v0x = ~p5 & ~p4 & ~p3 & ~p2;                        // 00..03
    const uint32_t v3x = p5 & p4 & p3;                                 // 38..3F
    L.X4b = cont & ((v0x & ~(p1 & p0)) | (v3x & ((~p2 & p1 & ~p0) | (p2 & ~p1 & ~p0))));
    L.S4b = cont & ((v0x & p1 & p0) | (v3x & ((~p2 & p1 & p0) | (p2 & p1))));
    {
        const uint32_t ge21 = p5 | (p4 & (p3 | (p2 & (p1 | p0))));
        const uint32_t ge58 = p5 & p4 & p3 & (p2 | p1);
        L.C4b = cont & ge21 & ~ge58;
    }
    L.M4b = cont & ((v0x & ~(p1 & p0)) | (v3x & p2 & p1));
    // after A5 (U+0940 + v): SpacingMark 00 09-0C 0E 0F, Extend 01-08 0D 11-17 22 23, linker 0D, consonant 18-1F 38-3F,
    // matra 00-0D 11-14
    const uint32_t v00_0f = ~p5 & ~p4, v10_1f = ~p5 & p4;
    const uint32_t sm5 = v00_0f & ((~p3 & lo3z) | (p3 & ((~p2 & (p1 | p0)) | (p2 & ~p1 & ~p0) | (p2 & p1))));
    L.S5b = cont & sm5;
    const uint32_t x5a = v00_0f & ~sm5;                                // 01-08, 0D  (everything in 00-0F that is not SM)
    const uint32_t x5b = v10_1f & ~p3 & (p2 | p1 | p0);                // 11-17
    const uint32_t x5c = p5 & ~p4 & ~p3 & ~p2 & p1;                    // 22 23
    L.X5b = cont & (x5a | x5b | x5c);
    L.LKb = cont & v00_0f & p3 & p2 & ~p1 & p0;                        // 0D
    L.C5b = cont & ((v10_1f & p3) | v3x);
    L.M5b = cont & ((v00_0f & ~(p3 & p2 & p1)) | (v10_1f & ~p3 & ((~p2 & (p1 | p0)) | (p2 & ~p1 & ~p0))));
    // ASCII
    const uint32_t row01 = asc & ~p6 & ~p5;                            // 00..1F
    L.CR = row01 & ~p4 & p3 & p2 & ~p1 & p0;
    L.LF = row01 & ~p4 & p3 & ~p2 & p1 & ~p0;
    const uint32_t del = asc & p6 & p5 & p4 & p3 & p2 & p1 & p0;
    L.CTL = (row01 | del) & ~(L.CR | L.LF);
    const uint32_t low_nz = p4 | p3 | p2 | p1 | p0;
    const uint32_t low_gt26 = p4 & p3 & (p2 | (p1 & p0));
    L.ROM = asc & p6 & low_nz & ~low_gt26;
    const uint32_t row2 = asc & ~p6 & p5 & ~p4, row3 = asc & ~p6 & p5 & p4;
    // punct: space ! " ' ( ) , - .   |   : ; ?   |   [ ] { }        digit: 30..39
    const uint32_t pk2 = (~p3 & ((~p2 & ~(p1 & p0)) | (p2 & p1 & p0))) | (p3 & ((~p2 & ~p1) | (p2 & ~(p1 & p0))));
    const uint32_t pk3 = p3 & ((~p2 & p1) | (p2 & p1 & p0));
    const uint32_t dg3 = ~p3 | (~p2 & ~p1);
    const uint32_t br = asc & p6 & p4 & p3 & ((~p2 & p1 & p0) | (p2 & ~p1 & p0));   // x1011 / x1101 in rows 5 and 7: [ ] { }
    L.WEAK = (row2 & pk2) | (row3 & (pk3 | dg3)) | br;
    L.dn1 = (L.A4b & 1u) | ((L.A5b & 1u) << 1) | ((L.X4b & 3u) << 2) | ((L.S4b & 3u) << 4) | ((L.C4b & 3u) << 6) |
            ((L.X5b & 3u) << 8) | ((L.S5b & 3u) << 10) | ((L.C5b & 3u) << 12) | ((L.LKb & 3u) << 14) | ((L.M4b & 3u) << 16) |
            ((L.M5b & 3u) << 18);
}

// leads reachable from the marked leads M through code points whose bytes are all in XB (the marked ones included)
AK_HD uint32_t aks3_star(uint32_t M, uint32_t XB) { return (((M & XB) + XB) ^ XB) | M; }

// ---- phase 2: code-point classes, end-of-lane summary (assuming nothing comes in from the left) -----------------
AK_HD void aks3_phase2(AkS3Lane& L, uint32_t dn1n) {
    const uint32_t a4_1 = akb_fsr(L.A4b, dn1n, 1), a5_1 = akb_fsr(L.A5b, dn1n >> 1, 1);
    const uint32_t d4 = L.E0b & a4_1, d5 = L.E0b & a5_1;
    L.DEV = d4 | d5;
    L.FOR = L.hl & ~L.DEV;
    const uint32_t x = (d4 & akb_fsr(L.X4b, dn1n >> 2, 2)) | (d5 & akb_fsr(L.X5b, dn1n >> 8, 2));
    L.X = x;
    L.SM = (d4 & akb_fsr(L.S4b, dn1n >> 4, 2)) | (d5 & akb_fsr(L.S5b, dn1n >> 10, 2));
    L.CONS = (d4 & akb_fsr(L.C4b, dn1n >> 6, 2)) | (d5 & akb_fsr(L.C5b, dn1n >> 12, 2));
    L.LK = d5 & akb_fsr(L.LKb, dn1n >> 14, 2);
    L.MAT = (d4 & akb_fsr(L.M4b, dn1n >> 16, 2)) | (d5 & akb_fsr(L.M5b, dn1n >> 18, 2));
    L.XI = x;
    L.XE = x;
    L.ZW = L.EP = L.RI = L.UNS = 0;
    const uint32_t ascl = L.lead & ~L.hl;
    L.strong_d = L.DEV;
    L.strong_r = L.ROM;
    L.strong_o = ascl & ~L.ROM & ~L.WEAK;
}

// code points outside ASCII / U+0900-097F, one by one from the property table (emoji, ZWJ, variation selectors, accents,
// other scripts when the text was not normalized first): they join the same masks, so the rules below hold for them too
AK_HD void aks3_foreign(const AkTables& Tb, const uint8_t* text, int64_t cs, int64_t te, AkS3Lane& L) {
    for (uint32_t m = L.FOR; m;) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        const uint32_t bit = 1u << i;
        int len;
        const uint32_t cp = ak_decode(text, cs + i, te, len);
        const uint32_t w = ak_props(Tb, cp);
        const uint32_t g = AK_GCB(w), ib = AK_INCB(w), tg = AK_TAG(w);
        if (g == GCB_EXTEND) { L.X |= bit; L.XE |= bit; }
        else if (g == GCB_ZWJ) { L.X |= bit; L.ZW |= bit; }
        else if (g == GCB_SPACINGMARK) L.SM |= bit;
        else if (g == GCB_CONTROL) L.CTL |= bit;
        else if (g == GCB_RI) L.RI |= bit;
        else if (g != GCB_OTHER) L.UNS |= bit;                  // Prepend, Hangul L / V / T / LV / LVT (CR / LF are ASCII)
        if (ib == INCB_CONSONANT) L.CONS |= bit;
        else if (ib == INCB_LINKER) { L.LK |= bit; L.XI |= bit; }
        else if (ib == INCB_EXTEND) L.XI |= bit;
        if (AK_EXTPICT(w)) L.EP |= bit;
        if (tg == TAG_DIGIT || tg == TAG_PUNCT) L.WEAK |= bit;
        else L.strong_o |= bit;                                  // identify_script: neither Devanagari nor A-Z a-z -> other
    }
}

// bytes of the code points led at M (lead + continuation bytes)
AK_HD uint32_t aks3_bytes_of(uint32_t M, uint32_t C) {
    uint32_t s = M;
    s |= (s << 1) & C;
    s |= (s << 1) & C;
    s |= (s << 1) & C;
    return s;
}

// conjunct state (GB9c) BEFORE the code point led at each position: r1 = "a consonant, then only Extend / Linker",
// r2 = "... with a Linker among them".  c1 / c2 = the state after the last code point before the lane.
AK_HD void aks3_conj(const AkS3Lane& L, uint32_t c1, uint32_t c2, uint32_t& r1, uint32_t& r2) {
    const uint32_t C = L.cont;
    const uint32_t bar = L.rows | ~L.own;
    const uint32_t fl = L.lead & (0u - L.lead);
    // bytes of the InCB Extend / Linker code points; the byte before a barrier is taken out so that no state crosses a
    // row start
    const uint32_t xb = aks3_bytes_of(L.XI, C) & ~(bar >> 1);
    const uint32_t t1 = (akb_fwd(L.CONS, C, 0u) | (c1 ? fl : 0u)) & ~bar;
    r1 = aks3_star(t1, xb) & L.lead;
    const uint32_t t2 = (akb_fwd(L.LK & r1, C, 0u) | (c2 ? fl : 0u)) & ~bar;
    r2 = aks3_star(t2, xb) & L.lead;
}
// GB11 state BEFORE each lead: p1 = "an ExtPict, then only Extend".  c1 = the state after the last code point before the lane.
AK_HD uint32_t aks3_pict(const AkS3Lane& L, uint32_t c1) {
    const uint32_t C = L.cont;
    const uint32_t bar = L.rows | ~L.own;
    const uint32_t fl = L.lead & (0u - L.lead);
    const uint32_t xb = aks3_bytes_of(L.XE, C) & ~(bar >> 1);
    const uint32_t t1 = (akb_fwd(L.EP, C, 0u) | (c1 ? fl : 0u)) & ~bar;
    return aks3_star(t1, xb) & L.lead;
}
// GB12 / GB13: leads with an ODD number of regional indicators right before them (in the same row).
// c_ri = the last code point before the lane is an RI, c_odd = an odd number of them ends there.
AK_HD uint32_t aks3_ri_odd(const AkS3Lane& L, uint32_t c_ri, uint32_t c_odd) {
    if (!L.RI && !c_odd) return 0u;
    const uint32_t C = L.cont;
    const uint32_t bar = L.rows | ~L.own;
    const uint32_t fl = L.lead & (0u - L.lead);
    const uint32_t prev_ri = akb_fwd(L.RI, C, c_ri) & ~bar;
    // pair starts: an RI with an even number of RIs before it -- the first of a run, or the lane's first lead when an even
    // number of RIs ends the previous lane
    const uint32_t starts = (L.RI & ~prev_ri) | ((c_ri && !c_odd) ? (L.RI & fl & ~bar) : 0u);
    uint32_t odd = (akb_fwd(starts, C, 0u) | (c_odd ? fl : 0u)) & ~bar;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t even = akb_fwd(odd & L.RI, C, 0u) & ~bar;
        odd |= akb_fwd(even & L.RI, C, 0u) & ~bar;
    }
    return odd;
}

// ---- phase 2b: end-of-lane summary for the next lane (computed with nothing coming in) --------------------------
AK_HD void aks3_summary(AkS3Lane& L) {
    const uint32_t bar = L.rows | ~L.own;
    uint32_t up = 0;
    uint32_t r1, r2;
    aks3_conj(L, 0u, 0u, r1, r2);
    const uint32_t lead = L.lead;
    if (lead) {
        const uint32_t ll = 0x80000000u >> akb_clz(lead);
        if ((L.CTL | L.CR | L.LF) & ll) up |= AKS3_LAST_CTL;
        if (L.CR & ll) up |= AKS3_LAST_CR;
        if (L.MAT & ll) up |= AKS3_LAST_MAT;
        if ((L.CONS | (L.XI & r1)) & ll) up |= AKS3_S1;
        if (((L.LK & r1) | (L.XI & r2)) & ll) up |= AKS3_S2;
        if (L.UNS & ll) up |= AKS3_LAST_UNS;
        if (L.EP | L.ZW) {
            const uint32_t p1 = aks3_pict(L, 0u);
            if ((L.EP | (L.XE & p1)) & ll) up |= AKS3_P1;
            if (L.ZW & p1 & ll) up |= AKS3_P2;
        }
        if (L.RI & ll) {
            up |= AKS3_LAST_RI;
            if (!(aks3_ri_odd(L, 0u, 0u) & ll)) up |= AKS3_RI_ODD;       // even before it -> odd with it
        }
    }
    // grapheme state: fixed by the last lead that is neither transparent for one of the rules (Extend, ZWJ, InCB Extend /
    // Linker, RI) nor unsupported, or by a barrier
    {
        const uint32_t sync = (lead & ~L.X & ~L.XI & ~L.RI & ~L.UNS) | bar;
        if (!sync) up |= AKS3_G_OPEN;
        else if (L.UNS && akb_clz(L.UNS) <= akb_clz(sync)) up |= AKS3_G_OPEN;      // (a barrier can sit on such a lead)
    }
    // run label at the lane's end: class of the last strong code point, none when a barrier follows it, open when the
    // lane has neither (or ends in unsupported code points)
    {
        const uint32_t strong = L.strong_d | L.strong_r | L.strong_o;
        const uint32_t known = strong | bar;
        uint32_t lab = 7u;
        if (known) {
            const uint32_t top = 0x80000000u >> akb_clz(known);
            lab = (L.strong_d & top) ? 1u : (L.strong_r & top) ? 2u : (L.strong_o & top) ? 3u : 0u;
            if (L.UNS && akb_clz(L.UNS) <= akb_clz(known)) lab = 7u;
        }
        up |= lab << AKS3_LAB_SHIFT;
    }
    L.up2 = up;
}

// ---- phase 3: boundaries and run changes given the previous lane's summary; false = slow lane -------------------
AK_HD bool aks3_phase3(AkS3Lane& L, uint32_t up2p, uint32_t tb_bit, bool matras, bool want_c, bool want_r) {
    const uint32_t C = L.cont;
    const uint32_t bar = L.rows | ~L.own;
    const uint32_t lead = L.lead;
    if (L.UNS) return false;
    if (!lead && !L.rows) { L.brk = L.rchg = L.PD = L.PR = L.PO = 0; return true; }
    const uint32_t fl = lead & (0u - lead);
    const bool first_at_bar = (fl & bar) != 0u || lead == 0u;
    if (want_c) {
        if (!first_at_bar && (up2p & (AKS3_LAST_UNS | AKS3_G_OPEN))) return false;
        uint32_t r1, r2;
        aks3_conj(L, (up2p & AKS3_S1) ? 1u : 0u, (up2p & AKS3_S2) ? 1u : 0u, r1, r2);
        const uint32_t p_ctl = akb_fwd(L.CTL | L.CR | L.LF, C, (up2p & AKS3_LAST_CTL) ? 1u : 0u);
        const uint32_t p_cr = akb_fwd(L.CR, C, (up2p & AKS3_LAST_CR) ? 1u : 0u);
        uint32_t nobreak = ((L.X | L.SM) & ~p_ctl) | (L.LF & p_cr) | (L.CONS & r2);
        if (L.EP) {                                                          // GB11: ExtPict Extend* ZWJ x ExtPict
            const uint32_t p1 = aks3_pict(L, (up2p & AKS3_P1) ? 1u : 0u);
            nobreak |= L.EP & akb_fwd(L.ZW & p1, C, (up2p & AKS3_P2) ? 1u : 0u) & ~bar;
        }
        if (L.RI)                                                            // GB12 / GB13: RI pairs
            nobreak |= L.RI & aks3_ri_odd(L, (up2p & AKS3_LAST_RI) ? 1u : 0u, (up2p & AKS3_RI_ODD) ? 1u : 0u);
        uint32_t brk = lead & ~nobreak;
        if (matras) brk |= (L.MAT | akb_fwd(L.MAT, C, (up2p & AKS3_LAST_MAT) ? 1u : 0u)) & lead;
        L.brk = brk & ~L.rows;
    } else L.brk = 0;
    if (want_r) {
        const uint32_t strong = L.strong_d | L.strong_r | L.strong_o;
        const uint32_t stops = strong | bar;
        const uint32_t lab = (up2p >> AKS3_LAB_SHIFT) & 7u;
        // the label in effect before the lane is not known, and something in the lane needs it
        if (lab == 7u && stops && !((stops & (0u - stops)) & tb_bit)) return false;
        const uint32_t skip = ~stops;
        uint32_t pd, pr, po;
        {
            uint32_t t = (L.strong_d << 1) | (lab == 1u ? 1u : 0u);
            pd = (((t & skip) + skip) | t) & stops;
            t = (L.strong_r << 1) | (lab == 2u ? 1u : 0u);
            pr = (((t & skip) + skip) | t) & stops;
            t = (L.strong_o << 1) | (lab == 3u ? 1u : 0u);
            po = (((t & skip) + skip) | t) & stops;
        }
        L.PD = pd; L.PR = pr; L.PO = po;
        L.rchg = ((L.strong_d & (pr | po)) | (L.strong_r & (pd | po)) | (L.strong_o & (pd | pr))) & ~L.rows;
    } else L.rchg = L.PD = L.PR = L.PO = 0;
    return true;
}

AK_HD uint32_t aks3_tag_at(const AkS3Lane& L, int i) {
    const uint32_t b = 1u << i;
    return (L.PD & b) ? (uint32_t)TAG_DEVANAGARI : (L.PR & b) ? (uint32_t)TAG_ROMAN : (L.PO & b) ? (uint32_t)TAG_OTHER : (uint32_t)TAG_NONE;
}

// Emission of one stream for a fast lane: every event (an in-row boundary, or a row start that closes a non-empty row)
// becomes its offset from the start of the row it ends in -- the highest row start strictly below it in the lane, else
// rs_in, the start of the row that was open when the lane began.  No row bookkeeping in the loop; the row splits are
// written separately (aks3_splits) from popcounts of the same masks.
AK_HD int aks3_emit(const AkS3Lane& L, uint32_t m, int64_t cs, int64_t rs_in, int32_t* dst, uint8_t* tags) {
    int k = 0;
    const uint32_t rows = L.rows;
    const int d0 = (int)(cs - rs_in);                       // 32-bit arithmetic in the loop: offsets are int32 by contract
    while (m) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        const uint32_t below = rows & ((1u << i) - 1u);
        dst[k] = below ? i - (31 - akb_clz(below)) : d0 + i;
        if (tags) tags[k] = (uint8_t)aks3_tag_at(L, i);
        ++k;
    }
    return k;
}

// splits of the rows that start in the lane: events of each stream at or before the row's position (+ the lane's base).
// nr = index of the first row that starts at or after the lane's first position; returns the next row index.
AK_HD int64_t aks3_splits(const AkS3Lane& L, uint32_t mc, uint32_t mr, int64_t cs, const int64_t* off, int64_t n_rows, int64_t nr,
                          int64_t cbase, int64_t rbase, int64_t* csplits, int64_t* rsplits) {
    for (uint32_t m = L.rows; m;) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        const int64_t p = cs + i;
        const uint32_t upto = (2u << i) - 1u;
        const int64_t kc = cbase + akb_popc(mc & upto), kr = rbase + akb_popc(mr & upto);
        while (nr <= n_rows && off[nr] == p) {
            if (csplits) csplits[nr] = kc;
            if (rsplits) rsplits[nr] = kr;
            ++nr;
        }
    }
    return nr;
}
