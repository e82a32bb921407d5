// Subword encoders of the hot path: the per-span BPE walker and the per-row
// Unigram Viterbi.  Like ak_text_core.cuh everything is AK_HD so the logic can also be compiled by g++ for the
// CPU walker tests; the product only runs the CUDA build.
//
// BPE (reference tokenizer.py:193 -> HuggingFace tokenizers; model JSON written by scripts/train_bpe.py:68-98):
//   NFKC (== NFC + exotic spaces on normalize_text's closed alphabet) -> words = maximal runs of \w or of
//   [^\w\s] -> per word: code points to single-character ids (characters outside the vocab are dropped, the
//   model has unk_token = null) -> merge the adjacent pair of lowest rank, leftmost first, until none is in the
//   merge table -> <s> ids </s> per row.
// Unigram (reference tokenizer.py:191 -> SentencePiece; ModelProto written by scripts/train_spm.py:80-108):
//   collapse / strip U+0020, dummy prefix, U+0020 -> U+2581 -> float32 Viterbi over code-point positions, strict
//   `>` so the earliest start wins ties, UNK edge (min_score - 10) where no single-character piece matches ->
//   UNK edges become byte pieces.
#pragma once
#include "ak_text_core.cuh"

#define AK_BPE_DIRECT 0x0A00      // code points below this are looked up in a direct table
#define AK_BPE_LOCAL 48           // symbols per word kept in thread-local arrays; longer words use the pool
#define AK_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

struct AkBpeDev {
    const int32_t* cp_direct;             // [AK_BPE_DIRECT] single-character token id or -1
    const uint32_t* cp_keys;              // sorted code points >= AK_BPE_DIRECT that have a token
    const int32_t* cp_ids;
    int n_cp;
    const unsigned long long* mkeys;      // (left_id << 32) | right_id, AK_EMPTY_KEY when free
    const unsigned long long* mvals;      // (rank << 32) | merged_id
    uint32_t mbits;                       // table has 1 << mbits slots
    int32_t bos, eos;                     // -1: the post-processor adds none
    const uint8_t* sp_bytes;              // added (special) tokens, longest first: bytes, offsets [n_sp + 1], ids
    const uint16_t* sp_off;
    const int32_t* sp_ids;
    int n_sp;
};

// HF AddedVocabulary (scripts/train_bpe.py:80): index of the added token that starts at byte p of the raw text (the longest
// one: they are stored longest first), or -1.  Never reads at or beyond `end`.
AK_HD int ak_bpe_special_at(const AkBpeDev& M, const uint8_t* t, int64_t p, int64_t end) {
    for (int k = 0; k < M.n_sp; ++k) {
        const int lo = M.sp_off[k], n = M.sp_off[k + 1] - lo;
        if (p + n > end) continue;
        int i = 0;
        while (i < n && t[p + i] == M.sp_bytes[lo + i]) ++i;
        if (i == n) return k;
    }
    return -1;
}

AK_HD uint32_t ak_hash64(unsigned long long k, uint32_t bits) {
    return (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> (64 - bits));
}

AK_HD_NOINLINE unsigned long long ak_bpe_pair(const AkBpeDev& M, int32_t a, int32_t b) {
    const unsigned long long key = ((unsigned long long)(uint32_t)a << 32) | (uint32_t)b;
    uint32_t h = ak_hash64(key, M.mbits);
    const uint32_t mask = (1u << M.mbits) - 1u;
    for (;;) {
        unsigned long long k = M.mkeys[h];
        if (k == key) return M.mvals[h];
        if (k == AK_EMPTY_KEY) return AK_EMPTY_KEY;
        h = (h + 1) & mask;
    }
}

AK_HD int32_t ak_bpe_char(const AkBpeDev& M, uint32_t cp) {
    if (cp < AK_BPE_DIRECT) return M.cp_direct[cp];
    int i = ak_bsearch<uint32_t>(M.cp_keys, M.n_cp, cp);
    return i < 0 ? -1 : M.cp_ids[i];
}

// HF `Word::merge_all` on sym[0..n): returns the new length.  rk[i] caches (rank << 32 | merged id) of the pair
// (sym[i], sym[i + 1]).
AK_HD int ak_bpe_merge(const AkBpeDev& M, int32_t* sym, unsigned long long* rk, int n) {
    // (loops kept rolled: this is the cache-miss path of an instruction-cache-bound kernel)
#pragma unroll 1
    for (int i = 0; i + 1 < n; ++i) rk[i] = ak_bpe_pair(M, sym[i], sym[i + 1]);
    while (n > 1) {
        unsigned long long best = AK_EMPTY_KEY;
        int bi = -1;
#pragma unroll 1
        for (int i = 0; i + 1 < n; ++i)
            if (rk[i] < best) { best = rk[i]; bi = i; }
        if (bi < 0) break;
        sym[bi] = (int32_t)(uint32_t)best;
#pragma unroll 1
        for (int i = bi + 1; i + 1 < n; ++i) { sym[i] = sym[i + 1]; rk[i] = rk[i + 1]; }
        --n;
        rk[bi] = (bi + 1 < n) ? ak_bpe_pair(M, sym[bi], sym[bi + 1]) : AK_EMPTY_KEY;
        if (bi > 0) rk[bi - 1] = ak_bpe_pair(M, sym[bi - 1], sym[bi]);
    }
    return n;
}

// per-thread id sink: the first `cap` ids go to buf[i * stride] (shared-memory staging, conflict-free layout);
// in `direct` mode ids go straight to global memory at gout[gbase + i]
struct AkIdSink {
    int32_t* buf;
    int cap, stride;
    int cnt;
    bool direct;
    int32_t* gout;
    int64_t gbase, gcap;
};
AK_HD void ak_id_put(AkIdSink& s, int32_t id) {
    if (s.direct) {
        if (s.gbase + s.cnt < s.gcap) s.gout[s.gbase + s.cnt] = id;
    } else if (s.cnt < s.cap) {
        s.buf[(int64_t)s.cnt * s.stride] = id;
    }
    ++s.cnt;
}

// long-word scratch: bump allocation of 3 ints per symbol from a global pool
struct AkPool {
    int32_t* base;
    unsigned long long* used;     // in ints
    unsigned long long cap;
};

AK_HD unsigned long long ak_pool_take(const AkPool& P, unsigned long long n) {
#ifdef __CUDA_ARCH__
    return atomicAdd(P.used, n);
#else
    unsigned long long v = *P.used;
    *P.used += n;
    return v;
#endif
}

// one pre-tokenized word starting at p (class k, row end re): symbolise, merge, emit.  Returns the word end.
AK_HD_NOINLINE int64_t ak_bpe_word(const AkBpeDev& M, const AkTables& T, const uint8_t* t, int64_t p, int64_t re,
                                   uint32_t k, AkIdSink& sink, const AkPool& pool, uint32_t& status) {
    int32_t sym[AK_BPE_LOCAL];
    unsigned long long rk[AK_BPE_LOCAL];
    int n = 0;
    int64_t q = p;
#pragma unroll 1
    while (q < re) {
        int len;
        uint32_t cp = ak_decode(t, q, re, len);
        if (AK_HFCLASS(ak_props(T, cp)) != k) break;
        int32_t id = ak_bpe_char(M, cp);
        if (id >= 0) {
            if (n < AK_BPE_LOCAL) sym[n] = id;
            ++n;
        }
        q += len;
    }
    if (n <= AK_BPE_LOCAL) {
        n = ak_bpe_merge(M, sym, rk, n);
#pragma unroll 1
        for (int i = 0; i < n; ++i) ak_id_put(sink, sym[i]);
        return q;
    }
    // long word: symbols + pair cache live in the global pool (8-byte aligned: 2 ints of cache per symbol first)
    unsigned long long need = 3ull * (unsigned long long)n + 1ull;
    unsigned long long at = ak_pool_take(pool, need);
    if (at + need > pool.cap) { status |= AK_ST_WORD; return q; }
    at = (at + 1ull) & ~1ull;
    unsigned long long* grk = (unsigned long long*)(pool.base + at);
    int32_t* gsym = pool.base + at + 2ull * (unsigned long long)n;
    int m = 0;
    int64_t c = p;
    while (c < q) {
        int len;
        uint32_t cp = ak_decode(t, c, re, len);
        int32_t id = ak_bpe_char(M, cp);
        if (id >= 0) gsym[m++] = id;
        c += len;
    }
    m = ak_bpe_merge(M, gsym, grk, m);
    for (int i = 0; i < m; ++i) ak_id_put(sink, gsym[i]);
    return q;
}

// does NFC change the segment that contains the code point at p?  (cold path: called on the first troubled code
// point of a segment only); *seg_end receives the end of that segment
AK_HD_NOINLINE bool ak_segment_changes(const AkTables& T, const uint8_t* t, int64_t p, int64_t rs, int64_t re,
                                       int64_t limit, int64_t* seg_end, uint32_t& status) {
    int64_t h = ak_find_head(T, t, p, rs, re, limit, status);
    bool trouble;
    int64_t hend = ak_scan_segment(T, t, h, re, trouble, limit, status);
    *seg_end = hend;
    uint32_t buf[AK_MAXSEG];
    int n = ak_nfc_segment(T, t, h, hend, buf, status);
    int i = 0;
    int64_t q = h;
    while (q < hend) {
        int len;
        uint32_t cp = ak_decode(t, q, hend, len);
        if (i >= n || buf[i] != cp) return true;
        ++i;
        q += len;
    }
    return i != n;
}

// Walk span [s, e): emits, in stream order, </s> of the previous row + <s> at every row start the span owns, the
// ids of every word that STARTS in the span, and the closing </s> at the end of the batch.  id_splits[r] (when
// non-null) receives split_base + the span-relative stream index of row r's first id, for r in
// [row_first, row_last) -- the caller adds the span's global base afterwards.  `changed` is set when NFC would
// alter a segment with a code point in this span (the caller then re-runs on NFC'd text).
AK_HD_NOINLINE void ak_bpe_span(const AkBpeDev& M, const AkTables& T, const uint8_t* t, const int64_t* off,
                                int64_t n_rows, int64_t r_lo, int64_t r_hi, int64_t s, int64_t e, int64_t limit,
                                AkIdSink& sink, int64_t* id_splits, int64_t split_base, int64_t& row_first,
                                int64_t& row_last, const AkPool& pool, bool& changed, uint32_t& status,
                                bool prenormalized = false) {
    const int64_t total_end = off[n_rows];
    row_first = row_last = 0;
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
    row_first = row_last = nr;
    int64_t rs = (off[nr] == p) ? p : off[nr - 1];
    int64_t re = (off[nr] == p) ? p : off[nr];
    uint32_t prev_class = 2, prev_ccc = 0;
    int64_t checked_until = -1;
    if (p != rs && p < total_end) {
        int64_t q = ak_prev_start(t, p, rs);
        int len;
        uint32_t w = ak_props(T, ak_decode(t, q, re, len));
        prev_class = AK_HFCLASS(w);
        prev_ccc = AK_CCC(w);
    }
    for (;;) {
        if (p >= e) break;
        while (nr <= n_rows && off[nr] == p) {
            if (nr > 0 && M.eos >= 0) ak_id_put(sink, M.eos);
            if (id_splits) id_splits[nr] = split_base + sink.cnt;
            if (nr < n_rows && M.bos >= 0) ak_id_put(sink, M.bos);
            ++nr;
            row_last = nr;
            rs = p;
            prev_class = 2;
            prev_ccc = 0;
        }
        if (p >= total_end) break;
        re = off[nr];
        int len;
        uint32_t cp = ak_decode(t, p, re, len);
        uint32_t w = ak_props(T, cp);
        // text the caller has put through HF's NFKC itself (the row-fix path) is taken as it is
        if (!prenormalized && !AK_BPE_SAFE(w)) status |= AK_ST_ALPHABET;
        uint32_t cc = AK_CCC(w);
        if (!prenormalized && (AK_QC(w) != 0 || (cc != 0 && prev_ccc > cc)) && p >= checked_until) {
            if (ak_segment_changes(T, t, p, rs, re, limit, &checked_until, status)) changed = true;
        }
        prev_ccc = cc;
        uint32_t k = AK_HFCLASS(w);
        if (k != 2 && k != prev_class) ak_bpe_word(M, T, t, p, re, k, sink, pool, status);
        prev_class = k;
        p += len;
    }
}

// =================================================================================================
// Unigram
// =================================================================================================
#define AK_UNI_RING 64            // > longest piece in code points
#define AK_UNI_UNKBIT 0x80000000u

struct AkUniDev {
    const unsigned long long* tkeys;      // (node << 21) | cp, AK_EMPTY_KEY when free
    const unsigned long long* tvals;      // (child << 32) | (piece_id + 1), low word 0 when no piece ends here
    const unsigned long long* tkv;        // optional: the same table interleaved (key, value) -- one 16-byte load per probe
    uint32_t tbits;
    const float* score;                   // per piece id (USER_DEFINED already resolved to len * max - 0.1)
    const uint8_t* usable;                // per piece id: 1 = takes part in the lattice (NORMAL / USER_DEFINED)
    const int32_t* byte_id;               // [256] id of <0xNN> or -1
    int32_t unk_id;
    float unk_score;
    int flags;                            // 1 add_dummy_prefix, 2 remove_extra_whitespaces, 4 escape_whitespaces, 8 byte_fallback
};

AK_HD unsigned long long ak_uni_child(const AkUniDev& U, uint32_t node, uint32_t cp) {
    const unsigned long long key = ((unsigned long long)node << 21) | cp;
    uint32_t h = ak_hash64(key, U.tbits);
    const uint32_t mask = (1u << U.tbits) - 1u;
#ifdef __CUDA_ARCH__
    if (U.tkv) {
        for (;;) {
            const ulonglong2 e = *reinterpret_cast<const ulonglong2*>(U.tkv + 2u * (size_t)h);
            if (e.x == key) return e.y;
            if (e.x == AK_EMPTY_KEY) return AK_EMPTY_KEY;
            h = (h + 1) & mask;
        }
    }
#endif
    for (;;) {
        unsigned long long k = U.tkeys[h];
        if (k == key) return U.tvals[h];
        if (k == AK_EMPTY_KEY) return AK_EMPTY_KEY;
        h = (h + 1) & mask;
    }
}

// normalized code point at byte position q of the trimmed row [ts, te); advances q past it (and past the
// U+0020 run it stands for when remove_extra_whitespaces)
AK_HD uint32_t ak_uni_next(const AkUniDev& U, const uint8_t* t, int64_t& q, int64_t te) {
    int len;
    uint32_t cp;
    const uint32_t b0 = t[q];
    if (b0 < 0x80u) { cp = b0; len = 1; }                       // the two shapes that make up the corpus, straight-line:
    else if ((b0 & 0xF0u) == 0xE0u && q + 3 <= te) {            // ASCII and three-byte code points (every Indic block)
        cp = ((b0 & 0x0Fu) << 12) | (((uint32_t)t[q + 1] & 0x3Fu) << 6) | ((uint32_t)t[q + 2] & 0x3Fu);
        len = 3;
    } else cp = ak_decode(t, q, te, len);
    q += len;
    if (cp == 0x20u) {
        if (U.flags & 2) while (q < te && t[q] == 0x20u) ++q;
        if (U.flags & 4) cp = 0x2581u;
    }
    return cp;
}

// forward Viterbi of one row; back[1..n] receives the final back-pointer of each position:
//   UNK edge: AK_UNI_UNKBIT | code point;  piece edge: (len << 24) | piece id.  Returns n (0 for an empty row).
// mark_byte / mark_index: *mark_index receives the lattice position of the code point that starts at byte mark_byte.
AK_HD_NOINLINE int64_t ak_unigram_forward(const AkUniDev& U, const uint8_t* t, int64_t rs, int64_t re, uint32_t* back,
                                          int64_t mark_byte = -1, int64_t* mark_index = nullptr) {
    int64_t ts = rs, te = re;
    if (U.flags & 2) {
        while (ts < te && t[ts] == 0x20u) ++ts;
        // trailing whitespace is removed from the ESCAPED output (sentencepiece normalizer.cc): with
        // escape_whitespaces a literal U+2581 (E2 96 81) at the end of the row goes too
        for (;;) {
            if (te > ts && t[te - 1] == 0x20u) --te;
            else if ((U.flags & 4) && te - ts >= 3 && t[te - 3] == 0xE2u && t[te - 2] == 0x96u && t[te - 1] == 0x81u) te -= 3;
            else break;
        }
    }
    if (ts >= te) return 0;
    float best[AK_UNI_RING];
    uint32_t bk[AK_UNI_RING];
    for (int i = 0; i < AK_UNI_RING; ++i) { best[i] = -INFINITY; bk[i] = 0; }
    best[0] = 0.0f;
    const bool dummy = (U.flags & 1) != 0;
    const uint32_t dummy_cp = (U.flags & 4) ? 0x2581u : 0x20u;
    int64_t i = 0;
    int64_t cur = ts;                 // byte position of normalized position i (position 0 is the dummy when set)
    for (;;) {
        const bool at_dummy = dummy && i == 0;
        if (!at_dummy && cur >= te) break;
        if (!at_dummy && cur == mark_byte && mark_index) *mark_index = i;
        if (i > 0) back[i] = bk[i & (AK_UNI_RING - 1)];
        const float bi = best[i & (AK_UNI_RING - 1)];
        // the slot that position i + RING - 1 will use is free from now on
        best[(i + AK_UNI_RING - 1) & (AK_UNI_RING - 1)] = -INFINITY;
        int64_t q = cur;
        uint32_t first_cp = 0;
        bool single = false;
        if (bi != -INFINITY) {
            uint32_t node = 0;
            for (int k = 1; k < AK_UNI_RING - 1; ++k) {
                uint32_t cp;
                if (at_dummy && k == 1) cp = dummy_cp;
                else {
                    if (q >= te) break;
                    cp = ak_uni_next(U, t, q, te);
                }
                if (k == 1) first_cp = cp;
                unsigned long long v = ak_uni_child(U, node, cp);
                if (v == AK_EMPTY_KEY) break;
                node = (uint32_t)(v >> 32);
                uint32_t pid1 = (uint32_t)v;
                if (pid1) {
                    uint32_t pid = pid1 - 1u;
                    if (U.usable[pid]) {
                        // sentencepiece 0.2.1 unigram_model.cc: the piece score is promoted to double by a ?: whose other
                        // arm is a double expression; the candidate is a DOUBLE sum, compared as a double against the
                        // stored float and rounded to float only when stored.  Reproduced operation for operation.
                        const double cand = (double)U.score[pid] + (double)bi;
                        int slot = (int)((i + k) & (AK_UNI_RING - 1));
                        if (cand > (double)best[slot]) { best[slot] = (float)cand; bk[slot] = ((uint32_t)k << 24) | pid; }
                        if (k == 1) single = true;
                    }
                }
            }
            if (!single) {
                float cand = bi + U.unk_score;
                int slot = (int)((i + 1) & (AK_UNI_RING - 1));
                if (cand > best[slot]) { best[slot] = cand; bk[slot] = AK_UNI_UNKBIT | first_cp; }
            }
        }
        // advance the row cursor by one normalized position
        if (!at_dummy) {
            int64_t c2 = cur;
            ak_uni_next(U, t, c2, te);
            cur = c2;
        }
        ++i;
    }
    back[i] = bk[i & (AK_UNI_RING - 1)];
    return i;
}

// walk the back-pointers from n; out == nullptr: count ids; else write them so that the row's ids end at out_end
AK_HD int64_t ak_unigram_backtrack(const AkUniDev& U, const uint32_t* back, int64_t n, int32_t* out, int64_t out_end,
                                   int64_t out_cap) {
    int64_t cnt = 0;
    int64_t t = n;
    while (t > 0) {
        uint32_t b = back[t];
        if (b & AK_UNI_UNKBIT) {
            uint32_t cp = b & 0x1FFFFFu;
            if (U.flags & 8) {
                uint8_t enc[4];
                int m = ak_encode(cp, enc);
                for (int j = m - 1; j >= 0; --j) {
                    ++cnt;
                    if (out && out_end - cnt < out_cap) out[out_end - cnt] = U.byte_id[enc[j]];
                }
            } else {
                ++cnt;
                if (out && out_end - cnt < out_cap) out[out_end - cnt] = U.unk_id;
            }
            t -= 1;
        } else {
            ++cnt;
            if (out && out_end - cnt < out_cap) out[out_end - cnt] = (int32_t)(b & 0xFFFFFFu);
            t -= (int64_t)((b >> 24) & 0x7Fu);
        }
    }
    return cnt;
}
