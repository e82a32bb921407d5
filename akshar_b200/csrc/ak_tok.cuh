// The subword encoders as dense kernels over an EVENT STREAM (the AK_HD cores live here; the kernels that drive them are
// in ak_tok_kernels.cuh):
//
//   words     text -> events: one per row start (payload = index of the first row that starts at that byte) and one per
//             pre-tokenized word (payload = its byte length).  The front end is the bit-stream classifier (ak_bpe3.cuh for
//             HF's `Whitespace` pre-tokenizer, aku3_* below for SentencePiece's space-delimited words).  Every warp owns 960
//             text bytes and a fixed block of event SLOTS (in text order, unused slots stay empty): warps need nothing
//             from one another -- no barrier, no scan, no look-back -- and the text is read once.
//   row fix   the rows the fast path does not take (not in NFC; a literal U+2581; very long rows) are encoded whole by the
//             exact row encoder into a side pool, their word events are struck out.
//   resolve   one thread per slot: word -> word cache (ak_wordcache.cuh) -> a 64-bit RESOLVED record (its ids inline, or
//             where to find them); a miss runs the exact encoder of the model (BPE merge loop / Unigram word lattice) and
//             publishes the word.  No ordering, so a slow miss delays nobody.
//   check     (Unigram) segmented sums along the rows decide which cached word lattices need the exact Viterbi.
//   emit      id counts from the resolved records -> block scan + decoupled look-back -> ids written once, at their final
//             place, as int32 or uint16.  Everything slow happened before: the tiles of this kernel take the same time, which
//             is what an ordered look-back needs (a tile's place is known only when ALL its predecessors have counted).
//
// Reference: tokenizer.py:191-193 (`EncodeAsIds(norm)` / `Tokenizer.encode(norm).ids`).
#pragma once
#include "ak_nfkc.cuh"
#include "ak_bpe3.cuh"
#include "ak_subword.cuh"
#include "ak_wordcache.cuh"

// event = (pos, meta): pos = byte offset from text_begin, meta = (payload << 3) | kind
#define AKE_OTHER 0u        // word of pre-tokenizer class "other" ([^\w\s]+)
#define AKE_WORD 1u         // word of class \w (HF) / a SentencePiece word
#define AKE_ROW 2u          // payload = index of the (one) row that starts at pos
#define AKE_ROWS 3u         // several rows start at pos (all but the last are empty): payload = index of the first
#define AKE_DEAD 4u         // overridden (its row is encoded by the row-fix kernel)
#define AKE_LEN_MAX 0x1FFFFFFFu

struct AkEvent {
    uint32_t pos, meta;
};

// ---- events of one lane (32 text bytes), in position order ------------------------------------------------------
// rowsm / wstart: the lane's row starts / word starts; cw: class-word mask; bnd + nb1 + nb2: boundary masks of this lane and
// the next two (a word's end = the next boundary); nr = index of the first row that starts at or after the lane's first
// byte.  `scan_end(p, from)` is called for the (cold) words with no boundary within what the warp knows.
// first_row / nrows: index of the first row that starts in the lane and how many do (akn3_lane_rows2): when every row
// bit stands for one row the rows are numbered from there, the offsets are only read again when rows are empty.
template <class ScanEnd>
AK_HD int ake_lane_events(uint32_t rowsm, uint32_t wstart, uint32_t cw, uint32_t bnd, uint32_t nb1, uint32_t nb2, int tail_known,
                          int64_t cs, int64_t tb, const int64_t* off, int64_t n_rows, int64_t first_row, int nrows, int64_t at,
                          uint32_t* row_ev, AkEvent* dst, int64_t cap_left, ScanEnd scan_end) {
    // Two loops instead of one over (rowsm | wstart): a slot index is a popcount, so the (few) row events do not hold up the
    // word events of the other lanes in every iteration.  At one position the row event comes first.
    const bool simple = nrows == akb_popc(rowsm);
    int64_t nr = first_row;
    int kr = 0;
    for (uint32_t m = rowsm; m; ++kr) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        const int64_t p = cs + i;
        const int k = kr + akb_popc(wstart & ((1u << i) - 1u));
        uint32_t kind = AKE_ROW;
        if (!simple) {
            while (nr < n_rows && off[nr] < p) ++nr;
            if (nr < n_rows && off[nr + 1] == p) kind = AKE_ROWS;
        }
        if (k < cap_left) { dst[k].pos = (uint32_t)(p - tb); dst[k].meta = ((uint32_t)nr << 3) | kind; }
        if (simple) row_ev[nr++] = (uint32_t)(at + k);
        else while (nr <= n_rows && off[nr] == p) row_ev[nr++] = (uint32_t)(at + k);      // empty rows share the position: one event
    }
    int kw = 0;
    for (uint32_t m = wstart; m; ++kw) {
        const int i = akb_ctz(m);
        m &= m - 1u;
        const int64_t p = cs + i;
        const int k = kw + akb_popc(rowsm & ((2u << i) - 1u));
        const uint32_t above = bnd & ~((2u << i) - 1u);
        int64_t len;
        if (above) len = akb_ctz(above) - i;
        else if (nb1) len = 32 - i + akb_ctz(nb1);
        else if (nb2) len = 64 - i + akb_ctz(nb2);
        else len = scan_end(p, cs + tail_known) - p;
        if (len > (int64_t)AKE_LEN_MAX) len = AKE_LEN_MAX;
        if (k < cap_left) { dst[k].pos = (uint32_t)(p - tb); dst[k].meta = ((uint32_t)len << 3) | ((cw >> i) & 1u); }
    }
    return kr + kw;
}

// the rows that hold the marked bytes of a lane are flagged for the row-fix kernel
AK_HD void ake_flag_rows(const int64_t* off, int64_t n_rows, int64_t cs, uint32_t mask, uint8_t* row_flag) {
    while (mask) {
        const int i = akb_ctz(mask);
        mask &= mask - 1u;
        const int64_t g = ak_row_lower_bound(off, 0, n_rows, cs + i + 1) - 1;
        if (g >= 0) row_flag[g] = 1;
    }
}

// ---- SentencePiece front end: words = maximal runs of bytes other than U+0020 -----------------------------------
// (sentencepiece normalizer: runs of U+0020 collapse, leading / trailing ones go, every word gets a U+2581 in front;
// any other white space is an ordinary character.)  A literal U+2581 in the text behaves like a space that does not
// collapse, and SentencePiece strips it at the end of a row: such rows are left to the exact row encoder (`exotic`).
struct AkU3Lane {
    uint32_t own, rows;
    uint32_t SP;          // 0x20 bytes
    uint32_t bnd, wstart;
    uint32_t exotic;      // E2 96 81 somewhere in the lane (or straddling into the next one: E2 96 at the end is flagged too)
    uint32_t up;          // bit 0: lane has an owned byte, bit 1: its last owned byte is a space
};
AK_HD void aku3_phase1(const uint32_t* x, AkU3Lane& L) {
    uint32_t P[8];
    akb_planes(x, P);
    const uint32_t p0 = P[0], p1 = P[1], p2 = P[2], p3 = P[3], p4 = P[4], p5 = P[5], p6 = P[6], p7 = P[7];
    L.SP = ~p7 & ~p6 & p5 & ~(p4 | p3 | p2 | p1 | p0) & L.own;
    const uint32_t e2 = p7 & p6 & p5 & ~p4 & ~p3 & ~p2 & p1 & ~p0;        // E2
    const uint32_t x96 = p7 & ~p6 & ~p5 & p4 & ~p3 & p2 & p1 & ~p0;       // 96
    // E2 followed by 96 (the third byte is not looked at: conservative), or E2 as the lane's last byte
    L.exotic = (e2 & ((x96 >> 1) | 0x80000000u)) & L.own;
    const uint32_t last = L.own ? (0x80000000u >> akb_clz(L.own)) : 0u;
    L.up = (L.own ? 1u : 0u) | ((L.SP & last) ? 2u : 0u);
}
// upp = the previous lane's `up` (a lane outside the text reads "nothing before")
AK_HD void aku3_phase2(AkU3Lane& L, uint32_t upp) {
    const uint32_t nsp = ~L.SP & L.own;
    // previous byte of the row is a non-space byte -> this byte continues a word
    const uint32_t prev_nsp = ((nsp << 1) | ((upp & 1u) && !(upp & 2u) ? 1u : 0u)) & ~L.rows;
    const uint32_t prev_sp = ((L.SP << 1) | ((upp & 3u) == 3u ? 1u : 0u)) & ~L.rows;
    L.wstart = nsp & ~prev_nsp;
    // boundaries: word starts, and the first space after a word; row starts
    L.bnd = L.wstart | (L.SP & ~prev_sp) | L.rows;
}

// =================================================================================================================
// look-up side
// =================================================================================================================
struct AkTokModel {
    int kind;                 // 0 BPE, 1 Unigram
    AkBpeDev bpe;
    AkUniDev uni;
    AkWordCache cache;
    AkPool pool;              // BPE long-word scratch
    AkTables T;
};

// ---- Unigram: the lattice of ONE word (U+2581 + body), reference tokenizer.py:191 -> sentencepiece unigram_model.cc
// EncodeOptimized restricted to the word.  No piece of a model trained with split_by_whitespace holds U+2581 anywhere
// but in front (checked at load), so every path of the sentence lattice passes through every word start: the sentence's
// best path is the concatenation of the words' best paths, PROVIDED float rounding at the magnitude of the score
// accumulated before the word cannot flip a decision inside it.  The word is therefore solved once, in double, from a
// zero start, together with
//     ratio = (smallest gap between the best and the second-best candidate of any lattice node) / (code points + 1)
//     wmag  = largest |partial score| seen in the word's lattice
// and a use of the cached result is exact when  ratio > 2^-22 * (sum of wmag over the row's words up to this one):
// float rounding moves every partial sum by at most (code points) * ulp(max magnitude) / 2, twice that separates two
// candidates, and the bound is doubled once more.  A word that fails the test takes its ids from the exact Viterbi of
// its whole row (ak_unigram_forward, SentencePiece's own float / double arithmetic).
#define AKU_WORD_CPS 64           // longest word (in code points, U+2581 included) the word lattice handles

struct AkUniWord {
    int n_ids;
    int32_t ids[AKU_WORD_CPS * 4];     // byte fallback: up to 4 ids per code point
    float ratio, wmag;
    bool ok;                           // false: word too long for the word lattice
};

AK_HD_NOINLINE void aku_word_lattice(const AkUniDev& U, const uint8_t* t, int64_t s, uint32_t len, AkUniWord& W) {
    uint32_t cps[AKU_WORD_CPS];
    int n = 0;
    cps[n++] = (U.flags & 4) ? 0x2581u : 0x20u;
    W.ok = true;
    W.n_ids = 0;
    W.ratio = 0.f;
    W.wmag = 0.f;
    {
        int64_t q = s;
        const int64_t e = s + len;
        while (q < e) {
            int l;
            const uint32_t cp = ak_decode(t, q, e, l);
            if (n >= AKU_WORD_CPS) { W.ok = false; return; }
            cps[n++] = cp;
            q += l;
        }
    }
    double best[AKU_WORD_CPS + 1], second[AKU_WORD_CPS + 1];
    uint32_t bk[AKU_WORD_CPS + 1];
    for (int i = 0; i <= n; ++i) { best[i] = -INFINITY; second[i] = -INFINITY; bk[i] = 0; }
    best[0] = 0.0;
    double wmag = 0.0;
    for (int i = 0; i < n; ++i) {
        if (best[i] == -INFINITY) continue;
        bool single = false;
        uint32_t node = 0;
        for (int k = 1; i + k <= n && k < AK_UNI_RING - 1; ++k) {
            const unsigned long long v = ak_uni_child(U, node, cps[i + k - 1]);
            if (v == AK_EMPTY_KEY) break;
            node = (uint32_t)(v >> 32);
            const uint32_t pid1 = (uint32_t)v;
            if (!pid1) continue;
            const uint32_t pid = pid1 - 1u;
            if (!U.usable[pid]) continue;
            const int j = i + k;
            const double cand = best[i] + (double)U.score[pid];
            const double a = cand < 0 ? -cand : cand;
            if (a > wmag) wmag = a;
            if (cand > best[j]) { second[j] = best[j]; best[j] = cand; bk[j] = ((uint32_t)k << 24) | pid; }
            else if (cand > second[j]) second[j] = cand;
            if (k == 1) single = true;
        }
        if (!single) {
            const int j = i + 1;
            const double cand = best[i] + (double)U.unk_score;
            const double a = cand < 0 ? -cand : cand;
            if (a > wmag) wmag = a;
            if (cand > best[j]) { second[j] = best[j]; best[j] = cand; bk[j] = AK_UNI_UNKBIT | cps[i]; }
            else if (cand > second[j]) second[j] = cand;
        }
    }
    int cnt = 0;
    for (int j = n; j > 0;) {
        const uint32_t b = bk[j];
        if (b & AK_UNI_UNKBIT) { cnt += (U.flags & 8) ? ak_utf8_len(b & 0x1FFFFFu) : 1; j -= 1; }
        else { cnt += 1; j -= (int)((b >> 24) & 0x7Fu); }
    }
    W.n_ids = cnt;
    int at = cnt;
    for (int j = n; j > 0;) {
        const uint32_t b = bk[j];
        if (b & AK_UNI_UNKBIT) {
            const uint32_t cp = b & 0x1FFFFFu;
            if (U.flags & 8) {
                uint8_t enc[4];
                const int m = ak_encode(cp, enc);
                for (int q = m - 1; q >= 0; --q) W.ids[--at] = U.byte_id[enc[q]];
            } else W.ids[--at] = U.unk_id;
            j -= 1;
        } else {
            W.ids[--at] = (int32_t)(b & 0xFFFFFFu);
            j -= (int)((b >> 24) & 0x7Fu);
        }
    }
    double margin = INFINITY;
    for (int j = 1; j <= n; ++j)
        if (best[j] != -INFINITY && best[j] - second[j] < margin) margin = best[j] - second[j];
    const double r = margin / (double)n;
    W.ratio = r > 1e30 ? 1e30f : (float)r * 0.999f;         // rounded towards "less robust"
    W.wmag = (float)wmag * 1.001f + 1e-30f;
}

// =================================================================================================================
// per-event cores of the lookup kernel (AK_HD: the CPU harness runs the same code event by event)
// =================================================================================================================
AK_HD unsigned long long ak_atomic_add64(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    const unsigned long long o = *p;
    *p += v;
    return o;
#endif
}
AK_HD void ak_status_or(int64_t* result, uint32_t bits) {
    if (!bits) return;
#ifdef __CUDA_ARCH__
    atomicOr((unsigned long long*)&result[2], (unsigned long long)bits);
#else
    result[2] |= (int64_t)bits;
#endif
}

struct AkLookupCtx {
    AkTokModel M;
    const uint8_t* text;
    const int64_t* off;
    int64_t n_rows, tb, te;
    int64_t* result;                   // status bits are OR-ed into result[2]
    void* ids;
    int64_t id_cap;
    int ids_u16;
    void* splits;
    int splits_i32;
    const uint8_t* row_flag;           // rows encoded by the row-fix kernel
    const unsigned long long* row_fix; // per flagged row: (pool offset << 24) | id count
    int32_t* pool;                     // row-fix ids (read) + scratch of the look-up (bump allocated)
    unsigned long long* pool_used;
    unsigned long long pool_cap;
    int any_fix;
};

AK_HD void akl_put(const AkLookupCtx& X, int64_t at, int32_t id) {
    if (at < X.id_cap) {
        if (X.ids_u16) ((uint16_t*)X.ids)[at] = (uint16_t)id;
        else ((int32_t*)X.ids)[at] = id;
    }
}
AK_HD void akl_split(const AkLookupCtx& X, int64_t g, int64_t v) {
    if (X.splits_i32) ((int32_t*)X.splits)[g] = (int32_t)v;
    else ((int64_t*)X.splits)[g] = v;
}

// ids of a row event: </s> of the previous row, the split, <s> (+ the ids of a fixed row); write = false: count only
AK_HD int akl_row_event(const AkLookupCtx& X, int64_t p, int64_t g, bool write, int64_t base) {
    const int32_t bos = X.M.kind == 0 ? X.M.bpe.bos : -1, eos = X.M.kind == 0 ? X.M.bpe.eos : -1;
    int k = 0;
    while (g <= X.n_rows && X.off[g] == p) {
        if (g > 0 && eos >= 0) { if (write) akl_put(X, base + k, eos); ++k; }
        if (write) akl_split(X, g, base + k);
        if (g < X.n_rows && bos >= 0) { if (write) akl_put(X, base + k, bos); ++k; }
        if (X.any_fix && g < X.n_rows && X.row_flag[g]) {
            const unsigned long long f = X.row_fix[g];
            const int c = (int)(f & 0xFFFFFFull);
            if (write) {
                const int32_t* src = X.pool + (f >> 24);
                for (int i = 0; i < c; ++i) akl_put(X, base + k + i, src[i]);
            }
            k += c;
        }
        ++g;
    }
    return k;
}

// cold: a BPE word that is not in the cache: the exact merge loop; the word is published when it fits an entry.
// Returns the id count; up to two ids come back inline (ids01), more go to the pool (*pool_at), a word with more ids
// than a private list holds is encoded again by the write phase (*pool_at = -2).
struct AkMissOut {
    int n;
    long long slot;                 // -1: up to two ids inline (ids01); <= -3: the ids are in the pool at offset -3 - slot
    unsigned long long ids01;
    float ratio, wmag;
    uint32_t st;
};
AK_HD_NOINLINE AkMissOut akl_bpe_miss(const AkLookupCtx& X, int64_t p, uint32_t len, uint32_t kc, long long free_slot,
                                      unsigned long long want, bool cacheable) {
    AkMissOut o;
    o.n = 0;
    o.slot = -1;
    o.ids01 = 0ull;
    o.ratio = 3.0e38f;
    o.wmag = 0.f;
    o.st = 0;
    uint32_t st = 0;
    if (!cacheable) {
        // a word longer than an entry: its ids (at most one per byte) go straight to the pool
        const unsigned long long at = ak_atomic_add64(X.pool_used, (unsigned long long)len);
        if (at + (unsigned long long)len > X.pool_cap) { o.st = AK_ST_WORD; return o; }
        AkIdSink s;
        s.buf = nullptr; s.cap = 0; s.stride = 1; s.cnt = 0; s.direct = true;
        s.gout = X.pool + at; s.gbase = 0; s.gcap = len;
        ak_bpe_word(X.M.bpe, X.M.T, X.text, p, p + len, kc, s, X.M.pool, st);
        o.n = s.cnt;
        o.slot = -3 - (long long)at;
        o.st = st;
        return o;
    }
    int32_t tmp[AKC_MAXLEN + 2];
    AkIdSink local;
    local.buf = tmp; local.cap = AKC_MAXLEN + 2; local.stride = 1; local.cnt = 0; local.direct = false;
    local.gout = nullptr; local.gbase = 0; local.gcap = 0;
    ak_bpe_word(X.M.bpe, X.M.T, X.text, p, p + len, kc, local, X.M.pool, st);
    const int n = local.cnt;
    // (a word whose encoding ran out of scratch is not to be remembered: the cache outlives the call)
    if (st == 0 && n <= AKC_MAXTOK && free_slot >= 0) akc_insert(X.M.cache, free_slot, want, X.text, p, len, tmp, n, 0ull);
    if (n <= 2) {
        o.ids01 = (n > 0 ? (unsigned long long)(uint32_t)tmp[0] : 0ull) | (n > 1 ? (unsigned long long)(uint32_t)tmp[1] << 32 : 0ull);
    } else {
        const unsigned long long at = ak_atomic_add64(X.pool_used, (unsigned long long)n);
        if (at + (unsigned long long)n > X.pool_cap) { o.st = st | AK_ST_WORD; return o; }
        for (int i = 0; i < n; ++i) X.pool[at + i] = tmp[i];
        o.slot = -3 - (long long)at;
    }
    o.n = n;
    o.st = st;
    return o;
}

// cold: a Unigram word that is not in the cache: solve its lattice, publish it when it fits an entry; the ids of this
// occurrence go to the pool.  Returns the id count.
AK_HD AkMissOut akl_uni_finish(const AkLookupCtx& X, int64_t p, uint32_t len, long long free_slot, unsigned long long want, bool cacheable,
                               bool ok, int n_ids, const int32_t* ids, float ratio, float wmag) {
    AkMissOut o;
    o.n = 0;
    o.slot = -1;
    o.ids01 = 0ull;
    o.st = 0;
    if (!ok) {
        // too long for the word lattice: always taken from the exact row Viterbi
        o.ratio = 0.f;
        o.wmag = (float)(len + 1u) * (X.M.uni.unk_score < 0 ? -X.M.uni.unk_score : X.M.uni.unk_score);
        return o;
    }
    // the tag's (bf16, conservative) numbers are what later look-ups see: use the same ones now
    const unsigned long long aux = akc_aux(ratio, wmag);
    o.ratio = akc_ratio(aux);
    o.wmag = akc_wmag(aux);
    if (cacheable && n_ids <= AKC_MAXTOK && free_slot >= 0) akc_insert(X.M.cache, free_slot, want, X.text, p, len, ids, n_ids, aux);
    if (n_ids <= 2) {
        o.ids01 = (n_ids > 0 ? (unsigned long long)(uint32_t)ids[0] : 0ull) | (n_ids > 1 ? (unsigned long long)(uint32_t)ids[1] << 32 : 0ull);
        o.n = n_ids;
        return o;
    }
    const unsigned long long at = ak_atomic_add64(X.pool_used, (unsigned long long)n_ids + 1ull);
    if (at + (unsigned long long)n_ids > X.pool_cap) { ak_status_or(X.result, AK_ST_WORD); return o; }
    for (int i = 0; i < n_ids; ++i) X.pool[at + i] = ids[i];
    o.slot = -3 - (long long)at;
    o.n = n_ids;
    return o;
}
AK_HD_NOINLINE AkMissOut akl_uni_miss(const AkLookupCtx& X, int64_t p, uint32_t len, long long free_slot, unsigned long long want,
                                      bool cacheable) {
    AkUniWord W;
    aku_word_lattice(X.M.uni, X.text, p, len, W);
    return akl_uni_finish(X, p, len, free_slot, want, cacheable, W.ok, W.n_ids, W.ids, W.ratio, W.wmag);
}

// cold: a Unigram word whose cached segmentation may depend on the score accumulated before it: its ids from the exact
// Viterbi of the row (SentencePiece's own arithmetic) up to the word's end.  Ids go to the pool.
AK_HD_NOINLINE int akl_uni_exact(const AkLookupCtx& X, int64_t p, uint32_t len, long long* pool_at_out) {
    long long pool_at = -1;
    *pool_at_out = -1;
    const int64_t g = ak_row_lower_bound(X.off, 0, X.n_rows, p + 1) - 1;
    const int64_t rs = X.off[g];
    const int64_t we = p + len;
    const unsigned long long need = (unsigned long long)(we - rs) + 4ull;
    const unsigned long long at = ak_atomic_add64(X.pool_used, need);
    if (at + need > X.pool_cap) { ak_status_or(X.result, AK_ST_WORD); return 0; }
    uint32_t* back = (uint32_t*)(X.pool + at);
    // the row up to the end of the word: every decision up to there is final (the lattice is built left to right), and
    // the word ends in a byte that is no space, so nothing is trimmed off the prefix
    int64_t mark = -1;
    const int64_t n = ak_unigram_forward(X.M.uni, X.text, rs, we, back, p, &mark);
    if (mark < 1) { ak_status_or(X.result, AK_ST_PATHOLOGICAL); return 0; }
    const int64_t a = mark - 1;              // lattice node in front of the word's U+2581
    int cnt = 0;
    for (int64_t t = n; t > a;) {
        const uint32_t b = back[t];
        if (b & AK_UNI_UNKBIT) { cnt += (X.M.uni.flags & 8) ? ak_utf8_len(b & 0x1FFFFFu) : 1; t -= 1; }
        else { cnt += 1; t -= (int64_t)((b >> 24) & 0x7Fu); }
    }
    const unsigned long long o = ak_atomic_add64(X.pool_used, (unsigned long long)cnt + 1ull);
    if (o + (unsigned long long)cnt > X.pool_cap) { ak_status_or(X.result, AK_ST_WORD); return 0; }
    int k = cnt;
    for (int64_t t = n; t > a;) {
        const uint32_t b = back[t];
        if (b & AK_UNI_UNKBIT) {
            const uint32_t cp = b & 0x1FFFFFu;
            if (X.M.uni.flags & 8) {
                uint8_t enc[4];
                const int m = ak_encode(cp, enc);
                for (int q = m - 1; q >= 0; --q) X.pool[o + (--k)] = X.M.uni.byte_id[enc[q]];
            } else X.pool[o + (--k)] = X.M.uni.unk_id;
            t -= 1;
        } else {
            X.pool[o + (--k)] = (int32_t)(b & 0xFFFFFFu);
            t -= (int64_t)((b >> 24) & 0x7Fu);
        }
    }
    pool_at = (long long)o;
    *pool_at_out = pool_at;
    return cnt;
}

// ---- resolved record: what the resolve kernel leaves for the emit kernel, 64 bits per slot ------------------------------
//   [63:62] 0 inline : [61:60] n (0..2), [59:30] id 1, [29:0] id 0          (an empty slot is all zero: n = 0)
//           1 cache  : [61:56] n (3..14), [55:0] cache entry
//           2 pool   : [61:38] n, [37:0] offset of the ids in the pool
//           3 event  : [61:38] n, [37] several rows start here (or the row was fixed), [36:0] index of the (first) row: a row start
#define AKR_INLINE 0ull
#define AKR_CACHE 1ull
#define AKR_POOL 2ull
#define AKR_EVENT 3ull
AK_HD unsigned long long akr_inline(int n, unsigned long long ids01) {
    return ((unsigned long long)n << 60) | (((ids01 >> 32) & 0x3FFFFFFFull) << 30) | (ids01 & 0x3FFFFFFFull);
}
AK_HD unsigned long long akr_cache(int n, long long slot) { return (AKR_CACHE << 62) | ((unsigned long long)n << 56) | (unsigned long long)slot; }
AK_HD unsigned long long akr_pool(int n, unsigned long long at) { return (AKR_POOL << 62) | ((unsigned long long)n << 38) | at; }
AK_HD unsigned long long akr_event(int n, int64_t g, bool multi) {
    return (AKR_EVENT << 62) | ((unsigned long long)n << 38) | (multi ? (1ull << 37) : 0ull) | (unsigned long long)g;
}
AK_HD int akr_n(unsigned long long r) {
    const unsigned long long ty = r >> 62;
    return ty == AKR_INLINE ? (int)((r >> 60) & 3ull) : ty == AKR_CACHE ? (int)((r >> 56) & 63ull) : (int)((r >> 38) & 0xFFFFFFull);
}

// resolve one event.  k[0..3]: the word's first four key words (akc_key0123) when it is a cacheable word.
// aux (Unigram): (ratio bf16 << 16) | wmag bf16 of the word, 0 for anything else.
template <int KIND>
AK_HD unsigned long long akl_resolve(const AkLookupCtx& X, AkEvent& ev, const unsigned long long* k, uint32_t& aux, uint32_t& st) {
    const uint32_t kind = ev.meta & 7u;
    uint32_t len = ev.meta >> 3;
    const int64_t p = X.tb + ev.pos;
    aux = 0u;
    if (kind <= AKE_WORD) {
        if (len == AKE_LEN_MAX) {
            // clamped in the event record (a word of half a gigabyte): find the real end again
            if (KIND == 0) len = (uint32_t)(akb3_scan_end(X.M.T, X.text, p, p + AKE_LEN_MAX, kind, X.off, X.n_rows, 0, X.n_rows) - p);
            ev.meta = (len << 3) | kind;
        }
        const bool cacheable = len <= AKC_MAXLEN;
        AkcHit h;
        h.slot = -1;
        h.free_slot = -1;
        h.h = h.want = h.tag = h.ids01 = 0ull;
        if (cacheable) akc_lookup4(X.M.cache, X.text, p, len, k, h);
        if (h.slot >= 0) {
            const int n = AKC_NTOK(h.tag);
            if (KIND == 1) aux = (uint32_t)(h.tag >> 32);
            return n <= 2 ? akr_inline(n, h.ids01) : akr_cache(n, h.slot);
        }
        const AkMissOut o = KIND == 0 ? akl_bpe_miss(X, p, len, kind, h.free_slot, h.want, cacheable)
                                      : akl_uni_miss(X, p, len, h.free_slot, h.want, cacheable);
        st |= o.st;
        if (KIND == 1) aux = (uint32_t)(akc_aux(o.ratio, o.wmag) >> 32);
        if (o.slot == -1) return akr_inline(o.n, o.ids01);
        return akr_pool(o.n, (unsigned long long)(-3 - o.slot));
    }
    if (kind == AKE_ROW) {
        // one row starts here: </s> of the previous row, <s> of this one (no look at the offsets)
        const int64_t g = (int64_t)(ev.meta >> 3);
        if (X.any_fix && g < X.n_rows && X.row_flag[g]) return akr_event(akl_row_event(X, p, g, false, 0), g, true);
        const int32_t bos = X.M.kind == 0 ? X.M.bpe.bos : -1, eos = X.M.kind == 0 ? X.M.bpe.eos : -1;
        return akr_event((g > 0 && eos >= 0 ? 1 : 0) + (g < X.n_rows && bos >= 0 ? 1 : 0), g, false);
    }
    if (kind == AKE_ROWS) return akr_event(akl_row_event(X, p, (int64_t)(ev.meta >> 3), false, 0), (int64_t)(ev.meta >> 3), true);
    return 0ull;
}

// Unigram: d = sum of wmag over the row's words up to and including this one; false = the cached segmentation is only
// trusted after the exact Viterbi has confirmed it (2^-22 * d, and 1 % on top for the float sums of d itself)
AK_HD bool aku_robust(float ratio, float d) { return ratio > d * (1.01f / 4194304.0f); }
AK_HD float aku_aux_ratio(uint32_t aux) {
    union { float f; uint32_t u; } a;
    a.u = aux & 0xFFFF0000u;
    return a.f;
}
AK_HD float aku_aux_wmag(uint32_t aux) {
    union { float f; uint32_t u; } a;
    a.u = aux << 16;
    return a.f;
}

// write the ids of one resolved slot at `at`
AK_HD_NOINLINE void akl_emit(const AkLookupCtx& X, unsigned long long r, const AkEvent* ev_slot, int64_t at) {
    const unsigned long long ty = r >> 62;
    // a record that points outside its table would be a bug of the resolve pass: say so instead of reading there
    if ((ty == AKR_CACHE && (r & 0xFFFFFFFFFFFFFFull) >= (1ull << X.M.cache.bits)) ||
        (ty == AKR_POOL && (r & 0x3FFFFFFFFFull) + ((r >> 38) & 0xFFFFFFull) > X.pool_cap) ||
        (ty == AKR_EVENT && (((ev_slot->meta & 7u) != AKE_ROW && (ev_slot->meta & 7u) != AKE_ROWS) || (int64_t)(ev_slot->meta >> 3) > X.n_rows))) {
        ak_status_or(X.result, AK_ST_INTERNAL | ((uint32_t)(ty + 1) << 8));
        return;
    }
    if (ty == AKR_INLINE) {
        const int n = (int)((r >> 60) & 3ull);
        if (n > 0) akl_put(X, at, (int32_t)(r & 0x3FFFFFFFull));
        if (n > 1) akl_put(X, at + 1, (int32_t)((r >> 30) & 0x3FFFFFFFull));
    } else if (ty == AKR_CACHE) {
        const int n = (int)((r >> 56) & 63ull);
        const unsigned long long* en = X.M.cache.e + (r & 0xFFFFFFFFFFFFFFull) * AKC_ENTRY;
        const unsigned long long i01 = akc_ld(en + 3);
        akl_put(X, at, (int32_t)(uint32_t)i01);
        akl_put(X, at + 1, (int32_t)(uint32_t)(i01 >> 32));
        for (int i = 2; i < n; ++i) akl_put(X, at + i, (int32_t)akc_id(en, i));
    } else if (ty == AKR_POOL) {
        const int n = (int)((r >> 38) & 0xFFFFFFull);
        const int32_t* src = X.pool + (r & 0x3FFFFFFFFFull);
        for (int i = 0; i < n; ++i) akl_put(X, at + i, src[i]);
    } else {
        const AkEvent ev = *ev_slot;
        const uint32_t kind = ev.meta & 7u;
        if (kind == AKE_ROW || kind == AKE_ROWS) akl_row_event(X, X.tb + ev.pos, (int64_t)(ev.meta >> 3), true, at);
    }
}

// =================================================================================================================
// row fix: one flagged row, encoded whole by the exact row encoder into the pool
// =================================================================================================================
struct AkRowFixCtx {
    AkTokModel M;
    const uint8_t* text;
    const int64_t* off;
    int64_t n_rows;
    int64_t* result;
    AkEvent* ev;
    unsigned long long n_events;
    const uint32_t* row_ev;
    unsigned long long* row_fix;
    int32_t* pool;
    unsigned long long* pool_used;
    unsigned long long pool_cap;
};

AK_HD_NOINLINE void akr_fix_row(const AkRowFixCtx& X, int64_t g) {
    const int64_t rs = X.off[g], re = X.off[g + 1];
    // strike the row's word events out (its start event stays and hands the ids over)
    {
        unsigned long long e0 = X.row_ev[g], e1 = X.row_ev[g + 1];
        if (e1 > X.n_events) e1 = X.n_events;
        for (unsigned long long e = e0 + 1; e < e1; ++e)
            if ((X.ev[e].meta & 7u) != AKE_ROW && (X.ev[e].meta & 7u) != AKE_ROWS) X.ev[e].meta = AKE_DEAD;
    }
    X.row_fix[g] = 0ull;
    if (X.M.kind == 1) {
        const unsigned long long need = (unsigned long long)(re - rs) + 4ull;        // back pointers: one per code point (+ dummy)
        const unsigned long long at = ak_atomic_add64(X.pool_used, need);
        if (at + need > X.pool_cap) { ak_status_or(X.result, AK_ST_WORD); return; }
        uint32_t* back = (uint32_t*)(X.pool + at);
        const int64_t n = ak_unigram_forward(X.M.uni, X.text, rs, re, back);
        const int64_t cnt = ak_unigram_backtrack(X.M.uni, back, n, nullptr, 0, 0);
        const unsigned long long at2 = ak_atomic_add64(X.pool_used, (unsigned long long)cnt + 1ull);
        if (at2 + (unsigned long long)cnt > X.pool_cap || cnt >= (1 << 24)) { ak_status_or(X.result, AK_ST_WORD); return; }
        ak_unigram_backtrack(X.M.uni, back, n, X.pool + at2, cnt, cnt);
        X.row_fix[g] = (at2 << 24) | (unsigned long long)cnt;
    } else {
        // BPE, the way HF runs it (scripts/train_bpe.py:68-98): the added tokens are cut out of the RAW row, every other
        // piece goes through NFKC, the Whitespace pre-tokenizer and the merges.  The normalized pieces are laid out in the
        // pool one after the other, each added token as the marker 0xFF + its index (no UTF-8 text holds 0xFF); the exact
        // walker then runs over each piece as a batch of one row.  <s> / </s> are the row event's business.
        uint32_t st = 0;
        const AkBpeDev& M = X.M.bpe;
        AkByteSink bs;
        bs.out = nullptr; bs.cnt = 0; bs.cap = 0;
        unsigned long long at = 0;
        for (int pass = 0; pass < 2; ++pass) {
            int64_t from = rs;
            for (int64_t q = rs; q <= re; ++q) {
                const int kk = q < re ? ak_bpe_special_at(M, X.text, q, re) : -1;
                if (q < re && kk < 0) continue;
                if (q > from) akk_nfkc(X.M.T, X.text, from, q, bs, st);
                if (q == re) break;
                if (bs.out && bs.cnt + 2 <= bs.cap) { bs.out[bs.cnt] = 0xFFu; bs.out[bs.cnt + 1] = (uint8_t)kk; }
                bs.cnt += 2;
                q += M.sp_off[kk + 1] - M.sp_off[kk] - 1;
                from = q + 1;
            }
            if (pass == 0) {
                const unsigned long long nints = (unsigned long long)(bs.cnt + 3) / 4ull + 2ull;
                at = ak_atomic_add64(X.pool_used, nints);
                if (at + nints > X.pool_cap) { ak_status_or(X.result, AK_ST_WORD | st); return; }
                bs.out = (uint8_t*)(X.pool + at);
                bs.cap = bs.cnt;
                bs.cnt = 0;
            }
        }
        const uint8_t* nt = (const uint8_t*)(X.pool + at);
        const int64_t nb = bs.cnt;
        AkBpeDev M2 = M;
        M2.bos = M2.eos = -1;
        AkIdSink sink;
        sink.buf = nullptr; sink.cap = 0; sink.stride = 1; sink.cnt = 0; sink.direct = false;
        sink.gout = nullptr; sink.gbase = 0; sink.gcap = 0;
        unsigned long long at2 = 0;
        for (int pass = 0; pass < 2; ++pass) {
            int64_t from = 0;
            for (int64_t q = 0; q <= nb; ++q) {
                if (q < nb && nt[q] != 0xFFu) continue;
                if (q > from) {
                    int64_t loff[2] = {from, q};
                    int64_t rf, rl;
                    bool changed = false;
                    ak_bpe_span(M2, X.M.T, nt, loff, 1, 0, 1, from, q + 1, 0, sink, nullptr, 0, rf, rl, X.M.pool, changed, st, true);
                }
                if (q == nb) break;
                ak_id_put(sink, M.sp_ids[nt[q + 1]]);
                ++q;
                from = q + 1;
            }
            if (pass == 0) {
                const int cnt = sink.cnt;
                at2 = ak_atomic_add64(X.pool_used, (unsigned long long)cnt + 1ull);
                if (at2 + (unsigned long long)cnt > X.pool_cap || cnt >= (1 << 24)) { ak_status_or(X.result, AK_ST_WORD | st); return; }
                sink.cnt = 0;
                sink.direct = true;
                sink.gout = X.pool + at2;
                sink.gbase = 0;
                sink.gcap = cnt;
            }
        }
        X.row_fix[g] = (at2 << 24) | (unsigned long long)sink.cnt;
        ak_status_or(X.result, st);
    }
}
