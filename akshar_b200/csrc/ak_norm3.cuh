// normalize_text (reference normalize.py:117-148, default flags) as parallel bit streams -- the classify step of
// the hot kernel, 32 text bytes per lane.
//
// The lane turns its bytes into basis planes (ak_bits.cuh) and derives, as 32-bit masks over its byte positions:
//   KL    lead bytes of the code points that survive filter_garbage (normalize.py:92-107): the allowed ASCII bytes
//         and every E0 A4/A5 xx (U+0900-097F); Bengali and everything else non-ASCII is resolved per code point by a
//         short table loop (AKN3 "loop" class) -- rare in the corpora this path is built for,
//   T     code points NFC (normalize.py:13-18) could touch: 0958-095F (QC=No), nukta not after an inert base,
//         virama / accent after an accent (ccc order), any non-plain code point found by the loop,
//   E     "same code point as the one right before it" (remove_elongations, normalize.py:48-56) from the plane
//         comparison with the bytes 1 / 3 positions earlier, A-Z compared as a-z (semantic_normalize runs first),
// and from them the emitted lead bytes: EM = KL & (~E | (~prevE & ~nextE)) -- a run of one or two survives, a
// longer run keeps its first element.  The output is exactly the emitted source bytes (A-Z lowered by the writer).
//
// Whatever these masks cannot decide exactly makes the lane SLOW (its two 16-byte chunks go to the walker's work
// list, ak_norm_span): an NFC segment with a T bit anywhere (also in the neighbouring lanes it reaches into), equal
// code points on both sides of a dropped stretch ("a<emoji>aa"), a pending run decision that a dropped stretch hides,
// kept non-ASCII other than U+0900-09FF, Latin capitals outside ASCII.  The fast lane is conservative, never
// approximate: tests/test_span_walkers.py runs these phases lane by lane on the CPU against the oracle.
//
// Cross-lane context (three exchange rounds, one packed word up and one down per round):
//   round 1  up: top 3 bits of every plane (for the byte comparisons)   down: low bits of the byte-role masks
//   round 2  up: INERT / accent / dropped at the last lead              down: Q3 low bits, first lead's E / dropped
//   round 3  up: E at the last lead, has-T, no-boundary                 down: T in the leading mark run, no-boundary
#pragma once
#include "ak_bits.cuh"
#include "ak_text_core.cuh"

struct AkN3Lane {
    uint32_t own;        // bit i: byte i lies inside the text
    uint32_t rows;       // bit i: a row starts at byte i
    // byte-role masks (phase 1)
    uint32_t cont, K, NL, E0b, F0b, LXo, A4b, A5b, A6b, A7b, x9Fb, NKb, NIb, R2b, R3b, QNb, B6b, B7b;
    uint32_t PL[8];      // basis planes, A-Z lowered
    // code-point masks at lead positions (phase 2 / 3)
    uint32_t lead, D3, NK, VIAC, AC, QNBT, INERT, KL, DROP, MARK, loopm, ge, Q1, Q3, E, T;
    uint32_t flags;
    // exchange words
    uint32_t up1, dn1, up2, dn2, up3, dn3;
};

#define AKN3_SLOW 1u

// ---- phase 1: planes and byte roles (no context) ---------------------------------------------------------------
AK_HD void akn3_phase1(const uint32_t* x, AkN3Lane& L) {
    uint32_t P[8];
    akb_planes(x, P);
    const uint32_t p0 = P[0], p1 = P[1], p2 = P[2], p3 = P[3], p4 = P[4], p5 = P[5], p6 = P[6], p7 = P[7];
    const uint32_t asc = ~p7;
    const uint32_t low_nz = p4 | p3 | p2 | p1 | p0;
    const uint32_t low_gt26 = p4 & p3 & (p2 | (p1 & p0));
    const uint32_t letter5 = low_nz & ~low_gt26;                      // low five bits in 1..26
    const uint32_t upper = asc & p6 & ~p5 & letter5;
    const uint32_t letters = asc & p6 & letter5;
    const uint32_t row2 = asc & ~p6 & p5 & ~p4;                       // 0x20..0x2F
    const uint32_t row3 = asc & ~p6 & p5 & p4;                        // 0x30..0x3F
    const uint32_t row0 = asc & ~p6 & ~p5 & ~p4;                      // 0x00..0x0F
    // 0x20 row: space ! " ' , - .      0x30 row: 0-9 : ; ?      0x00 row: 09..0D
    const uint32_t ok2 = (~p3 & ((~p2 & ~(p1 & p0)) | (p2 & p1 & p0))) | (p3 & p2 & ~(p1 & p0));
    const uint32_t ok3 = ~(p3 & p2) | (p1 & p0);
    const uint32_t ok0 = p3 & ((~p2 & (p1 | p0)) | (p2 & ~p1));
    L.K = (letters | (row2 & ok2) | (row3 & ok3) | (row0 & ok0)) & L.own;
    L.NL = row0 & p3 & ~p2 & p1 & ~p0;
    const uint32_t cont = p7 & ~p6;
    const uint32_t hl = p7 & p6;
    const uint32_t low4z = ~(p3 | p2 | p1 | p0);
    L.cont = cont;
    L.E0b = hl & p5 & ~p4 & low4z;
    L.F0b = hl & p5 & p4 & low4z;
    L.LXo = hl & ~(p5 & low4z);
    const uint32_t a4567 = cont & p5 & ~p4 & ~p3 & p2;               // 101001xx
    L.A4b = a4567 & ~p1 & ~p0;
    L.A5b = a4567 & ~p1 & p0;
    L.A6b = a4567 & p1 & ~p0;
    L.A7b = a4567 & p1 & p0;
    const uint32_t c01 = cont & ~p5;                                   // 100xxxxx
    L.x9Fb = c01 & p4 & p3 & p2 & p1 & p0;                            // 9F
    L.NKb = cont & p5 & p4 & p3 & p2 & ~p1 & ~p0;                     // BC   (after A4: U+093C)
    L.R2b = c01 & ~p4 & p3 & p2 & ~p1 & p0;                           // 8D   (after A5: U+094D)
    L.R3b = c01 & p4 & ~p3 & ((~p2 & (p1 | p0)) | (p2 & ~p1 & ~p0));  // 91..94 (after A5: U+0951-0954)
    L.QNb = c01 & p4 & p3;                                            // 98..9F (after A5: U+0958-095F)
    // A8 A9 B0 B1 B3 B4 (after A4: U+0928 0929 0930 0931 0933 0934 -- compose with / decompose to a nukta form)
    L.NIb = cont & p5 & ((~p4 & p3 & ~p2 & ~p1) | (p4 & ~p3 & ((~p2 & ~(p1 ^ p0)) | (~p2 & ~p1) | (p2 & ~p1 & ~p0))));
    // Bengali third bytes NFC cares about: BC BE after A6 (U+09BC 09BE); 87 8B 8C 8D 97 9C 9D 9F BE after A7
    L.B6b = cont & p5 & p4 & p3 & p2 & ~p0;
    {
        const uint32_t n7 = ~p3 & p2 & p1 & p0, cd = p3 & p2 & ~p1;
        L.B7b = (c01 & ~p4 & (n7 | (p3 & ~p2 & p1 & p0) | cd)) | (c01 & p4 & (n7 | cd | (p3 & p2 & p1 & p0))) |
                (cont & p5 & p4 & p3 & p2 & p1 & ~p0);
    }
    P[5] = p5 | upper;
    uint32_t up = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        L.PL[k] = P[k];
        up = (up << 3) | (P[k] >> 29);
    }
    L.up1 = up;
    L.dn1 = (cont & 7u) | ((L.A4b & 1u) << 3) | ((L.A5b & 1u) << 4) | ((L.x9Fb & 1u) << 5) | ((L.NKb & 3u) << 6) |
            ((L.NIb & 3u) << 8) | ((L.R2b & 3u) << 10) | ((L.R3b & 3u) << 12) | ((L.QNb & 3u) << 14) |
            ((L.A6b & 1u) << 16) | ((L.A7b & 1u) << 17) | ((L.B6b & 3u) << 18) | ((L.B7b & 3u) << 20);
}

// The warp's left halo lane has no left neighbour: it runs with up1 = 0 and the conservative up2 below (previous code
// point not inert, an accent, dropped).  Its own masks may then be off in its first bytes, but everything it hands to
// lane 1 is taken at its LAST leads; akn3_phase3b marks the hand-over unreliable in the rare lanes where those are the
// same code points (a last kept code point within the first three bytes, a blind gap end among the last two kept).
#define AKN3_HALO_UP2 6u

// ---- phase 2: byte comparisons, look-ahead roles, code-point classes ------------------------------------------
// up1p = the previous lane's up1, dn1n = the next lane's dn1
// raw = normalize_text(clean_hinglish=False): NFC + Roman lowercase only -- nothing is dropped, nothing collapses, so every
// lead is kept and E stays empty; the trouble logic is the same
AK_HD void akn3_phase2(AkN3Lane& L, uint32_t up1p, uint32_t dn1n, bool raw = false) {
    uint32_t q1 = 0xFFFFFFFFu, q3 = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t pw = up1p << (8 + 3 * k);                  // plane k's top three bits of the previous lane, at the top
        const uint32_t p = L.PL[k];
        q1 &= ~(p ^ akb_fsl(pw, p, 1));
        q3 &= ~(p ^ akb_fsl(pw, p, 3));
    }
    L.Q1 = q1;
    L.Q3 = q3;
    const uint32_t own = L.own;
    const uint32_t a4_1 = akb_fsr(L.A4b, dn1n >> 3, 1), a5_1 = akb_fsr(L.A5b, dn1n >> 4, 1), x9f_1 = akb_fsr(L.x9Fb, dn1n >> 5, 1);
    const uint32_t a6_1 = akb_fsr(L.A6b, dn1n >> 16, 1), a7_1 = akb_fsr(L.A7b, dn1n >> 17, 1);
    const uint32_t nk_2 = akb_fsr(L.NKb, dn1n >> 6, 2), ni_2 = akb_fsr(L.NIb, dn1n >> 8, 2), r2_2 = akb_fsr(L.R2b, dn1n >> 10, 2);
    const uint32_t r3_2 = akb_fsr(L.R3b, dn1n >> 12, 2), qn_2 = akb_fsr(L.QNb, dn1n >> 14, 2);
    const uint32_t b6_2 = akb_fsr(L.B6b, dn1n >> 18, 2), b7_2 = akb_fsr(L.B7b, dn1n >> 20, 2);
    const uint32_t e0 = L.E0b & own;
    const uint32_t d4 = e0 & a4_1, d5 = e0 & a5_1, d6 = e0 & a6_1, d7 = e0 & a7_1;
    L.lead = ~L.cont & own;
    L.D3 = d4 | d5 | d6 | d7;                                       // U+0900-09FF: all kept
    L.NK = d4 & nk_2;
    L.VIAC = d5 & (r2_2 | r3_2);
    L.AC = d5 & r3_2;
    L.QNBT = (d5 & qn_2) | (d6 & b6_2) | (d7 & b7_2);              // always troubled
    L.MARK = L.NK | L.VIAC | L.QNBT;
    L.INERT = (d4 | d5) & ~L.MARK & ~(d4 & ni_2);
    L.loopm = (L.LXo | (L.E0b & ~L.D3) | (L.F0b & ~x9f_1)) & own;
    L.KL = raw ? L.lead : (L.K | L.D3);
    L.DROP = L.lead & ~L.KL;
    L.up2 = akb_at_last(L.INERT, L.lead) | (akb_at_last(L.AC, L.lead) << 1) | (akb_at_last(L.DROP, L.lead) << 2);
    L.dn2 = q3 & 3u;
}

// the code point at `pos` as its (<= 4) bytes, A-Z lowered; never reads at or beyond te
AK_HD uint32_t akn3_cp_bytes(const uint8_t* t, int64_t pos, int64_t te) {
    uint32_t b0 = t[pos];
    if (b0 < 0x80u) return (b0 - 'A' < 26u) ? b0 + 32u : b0;
    int n = b0 >= 0xF0u ? 4 : b0 >= 0xE0u ? 3 : 2;
    uint32_t v = b0;
    for (int i = 1; i < n && pos + i < te; ++i) v |= (uint32_t)t[pos + i] << (8 * i);
    return v;
}

// ---- phase 3: E between adjacent code points, the loop class, trouble bits, gap ends ---------------------------
// up2p / dn2n: neighbours' words; cs = absolute position of byte 0
AK_HD void akn3_phase3(const AkTables& Tb, const uint8_t* text, int64_t cs, int64_t te, AkN3Lane& L, uint32_t up2p, uint32_t dn2n,
                       bool raw = false) {
    const uint32_t C = L.cont;
    {
        const uint32_t q3 = L.Q3;
        L.E = raw ? 0u : ((L.K & L.Q1) | (L.D3 & q3 & akb_fsr(q3, dn2n, 1) & akb_fsr(q3, dn2n, 2))) & ~L.rows & ~L.NL;
    }
    uint32_t flags = 0;
    uint32_t XT = 0;
    // loop class (neither ASCII nor U+0900-09FF nor U+1F000-1FFFF): decode and look up; all of them are dropped or slow
    for (uint32_t m = L.loopm; m;) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        int len;
        const uint32_t cp = ak_decode(text, cs + i, te, len);
        const uint32_t w = ak_props(Tb, cp);
        // not plain, or kept / turned into something kept (U+0130 -> i): the walker's business, and -- like a T bit --
        // the neighbours must not read this lane's kept code points as bytes
        if (!AK_NFC_HEAD(w) || AK_LATIN_LOWER(w) || (!raw && AK_ALLOW(w))) XT |= 1u << i;
    }
    const uint32_t p_inert = akb_fwd(L.INERT, C, up2p & 1u) & ~L.rows;
    const uint32_t p_ac = akb_fwd(L.AC, C, (up2p >> 1) & 1u) & ~L.rows;
    L.MARK |= XT;
    L.T = ((L.QNBT | (L.NK & ~p_inert) | (L.VIAC & p_ac)) & L.own) | XT;
    // kept code points right after a dropped stretch (same row): compared with the kept one before the stretch
    L.ge = raw ? 0u : (L.KL & akb_fwd(L.DROP, C, (up2p >> 2) & 1u) & ~L.rows & ~L.NL);
    L.flags = flags;
}

// gap ends whose previous kept code point lies in this lane: E from the byte comparison.  Returns the mask of the
// remaining ones (previous kept code point before the lane; at most the first gap end)
AK_HD uint32_t akn3_gaps_local(const uint8_t* text, int64_t cs, int64_t te, AkN3Lane& L) {
    uint32_t rest = 0;
    const uint32_t bar = L.rows | ~L.own;
    for (uint32_t m = L.ge; m;) {
        const int n = akb_ctz(m);
        m &= m - 1u;
        const uint32_t below = (1u << n) - 1u;
        const uint32_t kb = L.KL & below;
        const uint32_t rb = bar & ((below << 1) | 1u);           // barriers at or before n
        if (kb == 0) {
            if (rb == 0) rest |= 1u << n;
            continue;
        }
        const int p = 31 - akb_clz(kb);
        if ((rb >> p) >> 1) continue;                             // a barrier between p and n
        if (akn3_cp_bytes(text, cs + p, te) == akn3_cp_bytes(text, cs + n, te)) L.E |= 1u << n;
    }
    return rest;
}
// what a lane tells its successor about its last kept code point: 0 = none in this lane and no barrier either
// (unknown), 1 = a barrier after it / no kept one but a barrier (nothing to compare with), else its bytes
AK_HD uint32_t akn3_last_kept(const uint8_t* text, int64_t cs, int64_t te, const AkN3Lane& L) {
    const uint32_t bar = L.rows | ~L.own;
    if (L.KL == 0) return bar ? 1u : 0u;
    const int p = 31 - akb_clz(L.KL);
    if ((bar >> p) >> 1) return 1u;
    return akn3_cp_bytes(text, cs + p, te);
}
AK_HD void akn3_gaps_remote(const uint8_t* text, int64_t cs, int64_t te, AkN3Lane& L, uint32_t rest, uint32_t prev_last) {
    if (!rest) return;
    if (prev_last == 1u) return;
    if (prev_last == 0u) { L.flags |= AKN3_SLOW | 0x400u; L.ge = rest; return; }     // L.ge now holds the blind gap end
    const int n = akb_ctz(rest);
    if (akn3_cp_bytes(text, cs + n, te) == prev_last) L.E |= 1u << n;
}

// ---- phase 3b: summaries for round 3 ---------------------------------------------------------------------------
AK_HD void akn3_phase3b(AkN3Lane& L) {
    const uint32_t bar = L.rows | ~L.own;
    const uint32_t heads = (L.lead & ~L.MARK) | bar;
    const uint32_t below = heads ? ((heads & (0u - heads)) - 1u) : 0xFFFFFFFFu;
    const uint32_t no_boundary = heads ? 0u : 1u;
    const uint32_t first_dep = (L.MARK & below) ? 1u : 0u;
    const uint32_t lead_t = (L.T & below) ? 1u : 0u;
    const uint32_t has_t = L.T ? 1u : 0u;
    // tail trouble: a T bit at or after the head of the segment that holds the second-last kept code point --
    // what the next lane reads from this one (its last kept code point and that one's E) is then not to be trusted
    uint32_t tail_t = has_t;
    // a gap end whose other side nobody could see (L.ge then holds just that bit): it only matters to the next lane when it
    // is one of the last two kept code points
    const uint32_t blind = (L.flags & 0x400u) ? L.ge : 0u;
    uint32_t e_last = 0, unknown = 0;
    if (L.KL) {
        const int k1 = 31 - akb_clz(L.KL);
        if (!((bar >> k1) >> 1)) e_last = (L.E >> k1) & 1u;
        const uint32_t k2m = L.KL & ~(1u << k1);
        if (has_t && k2m) {
            const int k2 = 31 - akb_clz(k2m);
            const uint32_t hb = heads & ((2u << k2) - 1u);        // heads at or below k2
            if (hb) tail_t = (L.T >> (31 - akb_clz(hb))) ? 1u : 0u;
        }
        if (blind && (!k2m || (blind >> (31 - akb_clz(k2m))))) tail_t = 1u;
        // E at a lead in the first three bytes was compared with the previous lane's bytes: a lane without a left
        // neighbour (the warp's halo) has none, so a last kept code point that early is not to be trusted either
        if (k1 < 3) tail_t = 1u;
    } else if (!bar) unknown = 1u;
    if (blind && !L.KL) tail_t = 1u;
    // what the previous lane reads: E at the first kept code point -- not to be trusted when this lane has a T bit, or
    // when that code point's segment runs on to the lane's end (the trouble may sit in the lane after this one)
    uint32_t e_first = 0, head_bad = has_t;
    if (L.KL) {
        const int f = akb_ctz(L.KL);
        if (!(bar & ((1u << f) - 1u))) e_first = (L.E >> f) & 1u;
        if (((heads >> f) >> 1) == 0u) head_bad = 1u;
    }
    L.up3 = e_last | (tail_t << 1) | (no_boundary << 2) | (unknown << 3);
    L.dn3 = lead_t | (no_boundary << 1) | (head_bad << 2) | (e_first << 3) | (unknown << 4);
    if (L.T || no_boundary) L.flags |= AKN3_SLOW | 0x1000u;
    if (first_dep) L.flags |= 4u;
}

// ---- phase 4: emit masks of the two 16-byte chunks, or slow -----------------------------------------------------
// returns false when the lane is slow
AK_HD bool akn3_phase4(AkN3Lane& L, uint32_t up3p, uint32_t dn1n, uint32_t dn3n, uint32_t& info_lo, uint32_t& info_hi) {
    uint32_t flags = L.flags;
    if ((flags & 4u) && (up3p & 6u)) flags |= AKN3_SLOW | 0x2000u;   // leading marks of a segment that is troubled / headless before
    if (dn3n & 3u) flags |= AKN3_SLOW | 0x4000u;                     // my last segment reaches into trouble
    const uint32_t bar = L.rows | ~L.own;
    {
        // the previous lane's run state (its last kept code points as bytes) means nothing when NFC rewrites it
        const uint32_t before_row = bar ? ((bar & (0u - bar)) - 1u) : 0xFFFFFFFFu;
        if ((up3p & 2u) && (L.KL & before_row)) flags |= AKN3_SLOW | 0x8000u;
    }
    const uint32_t E = L.E;
    const uint32_t skip = ~L.KL & ~bar;                               // everything between two kept leads of one row
    const uint32_t pe = akb_fwd(E, skip, up3p & 1u) & L.KL & ~L.rows;
#ifdef __CUDA_ARCH__
    const uint32_t ne = __brev(akb_fwd(__brev(E), __brev(skip), (dn3n >> 3) & 1u)) & L.KL;
#else
    uint32_t ne;
    {
        uint32_t re = 0, rs = 0;
        for (int i = 0; i < 32; ++i) { re |= ((E >> i) & 1u) << (31 - i); rs |= ((skip >> i) & 1u) << (31 - i); }
        const uint32_t r = akb_fwd(re, rs, (dn3n >> 3) & 1u);
        ne = 0;
        for (int i = 0; i < 32; ++i) ne |= ((r >> i) & 1u) << (31 - i);
        ne &= L.KL;
    }
#endif
    if (L.KL) {
        // a pending second-of-run at the lane's end needs the next lane's first kept code point
        const int k1 = 31 - akb_clz(L.KL);
        if (((E & ~pe) >> k1) && !((bar >> k1) >> 1) && (dn3n & (4u | 16u))) flags |= AKN3_SLOW | 0x10000u;
    }
    L.flags = flags;
    if (flags & AKN3_SLOW) return false;
    const uint32_t em = L.KL & (~E | (~pe & ~ne));
    const uint32_t C = L.cont;
    const uint32_t chi = (C >> 16) | ((dn1n & 7u) << 16);
    uint32_t lo = em & 0xFFFFu, hi = em >> 16;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo |= (lo << 1) & C;
        hi |= (hi << 1) & chi;
    }
    info_lo = lo & 0x7FFFFu;
    info_hi = hi & 0x7FFFFu;
    return true;
}
