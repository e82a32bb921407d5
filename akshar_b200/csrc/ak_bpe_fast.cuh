// Fast lane of the BPE encoder (reference tokenizer.py:193 -> HF tokenizers `Tokenizer.encode(norm).ids`).
//
// Same work unit as the normalize fast lane (ak_fast.cuh): 16-byte chunks, 30 real + 2 halo chunks per warp.  A
// lane classifies its bytes (HF pre-tokenizer class per code point), finds the words that START in its chunk and
// encodes each one through a WORD CACHE in global memory (L2 resident): key = the word's bytes (<= 24, compared
// exactly), value = its token ids (<= 8).  HF's own BPE keeps the same kind of cache.  The cache image is built at
// model load from the vocabulary (every token string that is one pre-tokenizer word, encoded with the merge loop) and
// restored at the start of every call; words met during the call are added to it.  A miss runs the exact merge loop
// (ak_bpe_word) in place.  Row starts emit </s> <s> and the row split exactly like the walker ak_bpe_span.
#pragma once
#include "ak_fast.cuh"
#include "ak_subword.cuh"

#define AKW_MAXLEN 56          // key bytes per entry (7 x u64)
#define AKW_KW 7
#define AKW_MAXTOK 16          // ids per entry
#define AKW_ENTRY 16           // u64 per entry: tag, 7 key words, 8 x (2 ids)  = 128 bytes
#define AKW_PROBES 16         // linear probing; a look-up stops at the first empty slot, so long chains only cost the words that need them
#define AKW_READY 1ull
#define AKW_BUSY 2ull

struct AkWordCache {
    unsigned long long* e;      // (1 << bits) entries of AKW_ENTRY x u64
    uint32_t bits;
};

AK_HD unsigned long long akw_ld(const unsigned long long* p) {
#ifdef __CUDA_ARCH__
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#else
    return *p;
#endif
}
// look-up side: an ordinary (L1-cacheable) load.  Entries only ever go from empty to ready, the tag is published last
// behind a fence, and nobody reads an entry's other sectors before its tag matched, so a stale L1 sector can only show
// "still empty" -- a miss that is recomputed exactly -- never a torn entry.  The frequent words stay in L1.
AK_HD unsigned long long akw_ldc(const unsigned long long* p) {
#ifdef __CUDA_ARCH__
    unsigned long long v;
#ifdef AKW_PROBE_L2
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
#else
    asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
#endif
    return v;
#else
    return *p;
#endif
}
AK_HD void akw_st(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
    *p = v;
#endif
}

// Key word j (bytes [8j, 8j + 8) of the word [s, s + len), zero padded, little endian).  On the device: two 8-byte
// aligned loads + a funnel shift (the allocation of `t` must reach the next 8-byte boundary after its last byte, which
// holds for every CUDA allocation).  Everything below streams over the key words with ROLLED loops: this code sits on
// the hot path of an instruction-cache-bound kernel, compactness beats a few redundant L1 loads.
AK_HD unsigned long long akw_key_word(const uint8_t* t, int64_t s, uint32_t len, uint32_t j) {
    const uint32_t nb = len - 8u * j;           // bytes left from this word on (>= 1)
#ifdef __CUDA_ARCH__
    const uintptr_t a = (uintptr_t)(t + s) + 8u * j;
    const unsigned long long* base = (const unsigned long long*)(a & ~(uintptr_t)7);
    const uint32_t sh = (uint32_t)(a & 7u) * 8u;
    unsigned long long v = __ldg(base) >> sh;
    if (sh && (uint32_t)(a & 7u) + nb > 8u) v |= __ldg(base + 1) << (64u - sh);
#else
    unsigned long long v = 0ull;
    for (uint32_t i = 0; i < 8u && i < nb; ++i) v |= (unsigned long long)t[s + 8u * j + i] << (8u * i);
#endif
    if (nb < 8u) v &= (1ull << (8u * nb)) - 1ull;
    return v;
}

AK_HD unsigned long long akw_hash(const uint8_t* t, int64_t s, uint32_t len) {
    unsigned long long h = 0x9E3779B97F4A7C15ull + len;
    const uint32_t nw = (len + 7u) >> 3;
#pragma unroll 1
    for (uint32_t j = 0; j < nw; ++j) {
        h = (h ^ akw_key_word(t, s, len, j)) * 0xBF58476D1CE4E5B9ull;
        h = (h << 27) | (h >> 37);
    }
    h ^= h >> 31;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    return h;
}
// tag: [63:16] hash, [15:8] byte length, [7:3] token count, bit 1 busy, bit 0 ready
AK_HD unsigned long long akw_want(unsigned long long h, uint32_t len) { return (h & ~0xFFFFull) | ((unsigned long long)len << 8) | AKW_READY; }
#define AKW_NTOK_MASK 0xF8ull

// -> entry index of the word (its ids are then read by the caller), or -1; *free_slot = an empty slot seen on the probe
// path (or -1)
AK_HD long long akw_find(const AkWordCache& C, unsigned long long h, unsigned long long want, const uint8_t* t, int64_t s,
                         uint32_t len, long long* free_slot, unsigned long long* tag_out = nullptr) {
    const unsigned long long mask = (1ull << C.bits) - 1ull;
    const uint32_t nw = (len + 7u) >> 3;
    *free_slot = -1;
    const unsigned long long k0 = akw_key_word(t, s, len, 0);
    const unsigned long long k1 = nw > 1u ? akw_key_word(t, s, len, 1) : 0ull;
#pragma unroll 1
    for (int p = 0; p < AKW_PROBES; ++p) {
        const unsigned long long slot = (h + (unsigned long long)p) & mask;
        const unsigned long long* e = C.e + slot * AKW_ENTRY;
        // tag and the first two key words are fetched together (one sector): one round trip decides words up to 16 bytes
        const unsigned long long tag = akw_ldc(e);
        const unsigned long long e0 = akw_ldc(e + 1);
        const unsigned long long e1 = akw_ldc(e + 2);
        if (tag == 0ull) { *free_slot = (long long)slot; return -1; }
        if ((tag & ~AKW_NTOK_MASK) != want || e0 != k0 || (nw > 1u && e1 != k1)) continue;
        bool same = true;
#pragma unroll 1
        for (uint32_t j = 2; j < nw; ++j)
            if (akw_ldc(e + 1 + j) != akw_key_word(t, s, len, j)) { same = false; break; }
        if (same) {
            if (tag_out) *tag_out = tag;
            return (long long)slot;
        }
    }
    return -1;
}

// the first two key words of the word [s, s + len) without a branch: five aligned 32-bit loads (clamped to the last
// word of the text, whose bytes are masked off anyway) and four funnel shifts
AK_HD void akw_key01(const uint8_t* t, int64_t s, uint32_t len, int64_t te, unsigned long long& k0, unsigned long long& k1) {
#ifdef __CUDA_ARCH__
    const uintptr_t a = (uintptr_t)(t + s);
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t* last = (const uint32_t*)(((uintptr_t)(t + te) - 1u) & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3u) * 8u;
    const uint32_t w0 = __ldg(w);
    const uint32_t w1 = __ldg(w + 1 <= last ? w + 1 : last);
    const uint32_t w2 = __ldg(w + 2 <= last ? w + 2 : last);
    const uint32_t w3 = __ldg(w + 3 <= last ? w + 3 : last);
    const uint32_t w4 = __ldg(w + 4 <= last ? w + 4 : last);
    const uint32_t f0 = __funnelshift_r(w0, w1, sh), f1 = __funnelshift_r(w1, w2, sh);
    const uint32_t f2 = __funnelshift_r(w2, w3, sh), f3 = __funnelshift_r(w3, w4, sh);
    k0 = ((unsigned long long)f1 << 32) | f0;
    k1 = ((unsigned long long)f3 << 32) | f2;
    const unsigned long long m0 = len >= 8u ? ~0ull : ((1ull << (8u * len)) - 1ull);
    const unsigned long long m1 = len >= 16u ? ~0ull : (len > 8u ? ((1ull << (8u * (len - 8u))) - 1ull) : 0ull);
    k0 &= m0;
    k1 &= m1;
#else
    (void)te;
    k0 = akw_key_word(t, s, len, 0);
    k1 = len > 8u ? akw_key_word(t, s, len, 1) : 0ull;
#endif
}

// hash + probe with the first two key words computed once (they decide every word up to 16 bytes)
AK_HD long long akw_lookup(const AkWordCache& C, const uint8_t* t, int64_t s, uint32_t len, int64_t te, unsigned long long* h_out,
                           unsigned long long* want_out, long long* free_slot, unsigned long long* tag_out) {
    const uint32_t nw = (len + 7u) >> 3;
    unsigned long long k0, k1;
    akw_key01(t, s, len, te, k0, k1);
    unsigned long long h = 0x9E3779B97F4A7C15ull + len;
    h = (h ^ k0) * 0xBF58476D1CE4E5B9ull;
    h = (h << 27) | (h >> 37);
    if (nw > 1u) {
        h = (h ^ k1) * 0xBF58476D1CE4E5B9ull;
        h = (h << 27) | (h >> 37);
#pragma unroll 1
        for (uint32_t j = 2; j < nw; ++j) {
            h = (h ^ akw_key_word(t, s, len, j)) * 0xBF58476D1CE4E5B9ull;
            h = (h << 27) | (h >> 37);
        }
    }
    h ^= h >> 31;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    const unsigned long long want = akw_want(h, len);
    *h_out = h;
    *want_out = want;
    *free_slot = -1;
    const unsigned long long mask = (1ull << C.bits) - 1ull;
#pragma unroll 1
    for (int p = 0; p < AKW_PROBES; ++p) {
        const unsigned long long slot = (h + (unsigned long long)p) & mask;
        const unsigned long long* e = C.e + slot * AKW_ENTRY;
        const unsigned long long tag = akw_ldc(e);
        const unsigned long long e0 = akw_ldc(e + 1);
        const unsigned long long e1 = akw_ldc(e + 2);
        if (tag == 0ull) { *free_slot = (long long)slot; return -1; }
        if ((tag & ~AKW_NTOK_MASK) != want || e0 != k0 || (nw > 1u && e1 != k1)) continue;
        bool same = true;
#pragma unroll 1
        for (uint32_t j = 2; j < nw; ++j)
            if (akw_ldc(e + 1 + j) != akw_key_word(t, s, len, j)) { same = false; break; }
        if (same) {
            *tag_out = tag;
            return (long long)slot;
        }
    }
    return -1;
}

AK_HD void akw_insert(const AkWordCache& C, long long slot, unsigned long long want, const uint8_t* t, int64_t s, uint32_t len,
                      const int32_t* ids, int n) {
    unsigned long long* e = C.e + (unsigned long long)slot * AKW_ENTRY;
#ifdef __CUDA_ARCH__
    if (atomicCAS(e, 0ull, AKW_BUSY) != 0ull) return;
#else
    if (*e != 0ull) return;
    *e = AKW_BUSY;
#endif
    const uint32_t nw = (len + 7u) >> 3;
    for (uint32_t j = 0; j < nw; ++j) akw_st(e + 1 + j, akw_key_word(t, s, len, j));
    for (int i = 0; i < n; i += 2) {
        unsigned long long v = (uint32_t)ids[i];
        if (i + 1 < n) v |= (unsigned long long)(uint32_t)ids[i + 1] << 32;
        akw_st(e + 8 + (i >> 1), v);
    }
#ifdef __CUDA_ARCH__
    __threadfence();
#endif
    akw_st(e, want | ((unsigned long long)n << 3));
}

// encode the word [s, e) of pre-tokenizer class k into `sink`: cache hit, or the merge loop + insert
AK_HD_NOINLINE void akb_word(const AkBpeDev& M, const AkTables& T, const AkWordCache& C, const uint8_t* t, int64_t s,
                             int64_t e, uint32_t k, AkIdSink& sink, const AkPool& pool, uint32_t& status) {
    const uint32_t len = (uint32_t)(e - s);
    if (e - s <= AKW_MAXLEN && C.e) {
        const unsigned long long h = akw_hash(t, s, len);
        const unsigned long long want = akw_want(h, len);
        long long slot;
        const long long hit = akw_find(C, h, want, t, s, len, &slot);
        if (hit >= 0) {
            const unsigned long long* en = C.e + (unsigned long long)hit * AKW_ENTRY;
            const int n = (int)((akw_ld(en) & AKW_NTOK_MASK) >> 3);
#pragma unroll 1
            for (int i = 0; i < n; i += 2) {
                const unsigned long long v = akw_ld(en + 8 + (i >> 1));
                ak_id_put(sink, (int32_t)(uint32_t)v);
                if (i + 1 < n) ak_id_put(sink, (int32_t)(uint32_t)(v >> 32));
            }
            return;
        }
        // miss: exact merge loop into a private list, then publish
        int32_t tmp[AKW_MAXTOK + 1];
        AkIdSink local;
        local.buf = tmp;
        local.cap = AKW_MAXTOK + 1;
        local.stride = 1;
        local.cnt = 0;
        local.direct = false;
        local.gout = nullptr;
        local.gbase = 0;
        local.gcap = 0;
        ak_bpe_word(M, T, t, s, e, k, local, pool, status);
        if (local.cnt <= AKW_MAXTOK) {
            for (int i = 0; i < local.cnt; ++i) ak_id_put(sink, tmp[i]);
            if (slot >= 0) akw_insert(C, slot, want, t, s, len, tmp, local.cnt);
            return;
        }
    }
    ak_bpe_word(M, T, t, s, e, k, sink, pool, status);
}

// ------------------------------------------------------------------------------------------------
// per-chunk classification
// ------------------------------------------------------------------------------------------------
struct AkBChunk {
    uint32_t w[5];
    uint32_t rows, own;
    // phase A
    uint32_t lead;        // owned lead bytes
    uint32_t bnd;         // word boundaries: row starts, and owned leads whose class differs from the previous code point's
    uint32_t cls;         // 2 bits per byte position: HF pre-tokenizer class of the code point led there
    uint32_t last_cls;    // class of the last owned code point (2 = space when none)
    uint32_t last_w;      // its props (AKF_NONE when none)
    uint32_t first_pos;   // position of the first owned lead when it is not on a row start, else 32
    uint32_t first_w;     // its props
    uint32_t flags;       // AKF_TROUBLE (needs the NFC check), bit 8: outside the closed alphabet
    uint32_t trb;         // positions of the code points that raised AKF_TROUBLE
};
#define AKB_ALPHABET 256u

AK_HD uint32_t akb_byte(const AkBChunk& c, int i) { return (c.w[i >> 2] >> ((i & 3) * 8)) & 0xFFu; }

// The loop over the 16 byte positions is rolled over the 4 words (4 bytes unrolled inside): a fully unrolled body is
// ~4x the code, and these kernels are instruction-cache bound as soon as their warps drift apart.
AK_HD void akb_phase_a(const AkTables& T, const uint32_t* lut, AkBChunk& c) {
    uint32_t lead = 0, bnd = c.rows, cls = 0, flags = 0, trb = 0;
    uint32_t prev_w = AKF_NONE, prev_cls = 2;
    bool have_prev = false, first_seen = false;
    uint32_t first_pos = 32, first_w = AKF_NONE, last_w = AKF_NONE, last_cls = 2;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        // the word's 4 bytes and the 3 that follow, as one 64-bit window
        const uint32_t wlo = k == 0 ? c.w[0] : k == 1 ? c.w[1] : k == 2 ? c.w[2] : c.w[3];
        const uint32_t whi = k == 0 ? c.w[1] : k == 1 ? c.w[2] : k == 2 ? c.w[3] : c.w[4];
        const unsigned long long win = ((unsigned long long)whi << 32) | wlo;
        const uint32_t rows4 = (c.rows >> (4 * k)) & 15u, own4 = (c.own >> (4 * k)) & 15u;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int i = 4 * k + j;
            const bool row_here = (rows4 >> j) & 1u;
            if (row_here) { have_prev = false; prev_cls = 2; }
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
            const uint32_t kc = AK_HFCLASS(w);
            lead |= 1u << i;
            cls |= kc << (2 * i);
            if (!AK_BPE_SAFE(w)) flags |= AKB_ALPHABET;
            if (!first_seen && !row_here) {
                first_pos = (uint32_t)i;
                first_w = w;
                if (AK_QC(w) == 1u) { flags |= AKF_TROUBLE; trb |= 1u << i; }
            } else {
                if (!have_prev ? (AK_QC(w) != 0u) : akf_trouble_after(prev_w, w)) { flags |= AKF_TROUBLE; trb |= 1u << i; }
                if (row_here || kc != prev_cls) bnd |= 1u << i;
            }
            first_seen = true;
            prev_w = w;
            prev_cls = kc;
            have_prev = true;
            last_w = w;
            last_cls = kc;
        }
    }
    c.lead = lead;
    c.bnd = bnd;
    c.cls = cls;
    c.flags = flags;
    c.trb = trb;
    c.first_pos = first_pos;
    c.first_w = first_w;
    c.last_w = last_w;
    c.last_cls = last_cls;
}

// the first owned code point against the previous chunk's last one
AK_HD void akb_resolve_first(AkBChunk& c, uint32_t prev_last_w, uint32_t prev_last_cls) {
    if (c.first_pos >= 32u) return;
    if (prev_last_w == AKF_NONE ? (AK_QC(c.first_w) != 0u) : akf_trouble_after(prev_last_w, c.first_w)) {
        c.flags |= AKF_TROUBLE;
        c.trb |= 1u << c.first_pos;
    }
    const uint32_t k = (c.cls >> (2 * c.first_pos)) & 3u;
    if (prev_last_w == AKF_NONE || k != prev_last_cls) c.bnd |= 1u << c.first_pos;
}

// end of the word that starts at chunk byte s: next boundary in this chunk, else in the next one, else a forward scan
AK_HD int64_t akb_word_end(const AkTables& T, const uint8_t* t, int64_t cs, int s, uint32_t k, uint32_t bnd, uint32_t next_bnd,
                           const int64_t* off, int64_t n_rows, int64_t r_lo, int64_t r_hi) {
    const uint32_t above = bnd & ~((2u << s) - 1u) & 0xFFFFu;
    if (above) {
#ifdef __CUDA_ARCH__
        return cs + (__ffs(above) - 1);
#else
        return cs + __builtin_ctz(above);
#endif
    }
    if (next_bnd & 0xFFFFu) {
#ifdef __CUDA_ARCH__
        return cs + 16 + (__ffs(next_bnd & 0xFFFFu) - 1);
#else
        return cs + 16 + __builtin_ctz(next_bnd & 0xFFFFu);
#endif
    }
    if (next_bnd >> 16) {      // bits 16..31: the chunk after the next one (when the caller has it)
#ifdef __CUDA_ARCH__
        return cs + 16 + (__ffs(next_bnd) - 1);
#else
        return cs + 16 + __builtin_ctz(next_bnd);
#endif
    }
    // long word: walk code points up to the end of its row (no boundary before cs + 32); the row usually ends
    // inside the tile's row window, else search the whole offset array
    int64_t er = ak_row_lower_bound(off, r_lo, r_hi, cs + s + 1);
    if (off[er] < cs + s + 1) er = ak_row_lower_bound(off, r_hi, n_rows, cs + s + 1);
    const int64_t re = off[er];
    int64_t q = cs + 32;
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


// Everything one lane emits for its chunk, in stream order: </s> <s> + split at the row starts it holds, the ids of
// the words that start in it.  Mirrors ak_bpe_span.  id_splits[r] receives the lane-relative index for the rows
// [row_first, row_last) (the caller makes them global).
struct AkBLaneCtx {
    const AkBpeDev* M;
    const AkTables* T;
    const AkWordCache* C;
    const uint8_t* text;
    const int64_t* off;
    int64_t n_rows, r_lo, r_hi;
    const AkPool* pool;
};

AK_HD_NOINLINE void akb_lane_emit(const AkBLaneCtx& X, const AkBChunk& c, uint32_t next_bnd, int64_t cs, AkIdSink& sink,
                                  int64_t* id_splits, int64_t& row_first, int64_t& row_last, uint32_t& status,
                                  int64_t nr_hint = -1) {
    // word starts = boundaries at owned leads whose class is not "space" (2 = binary 10: high bit set, low bit clear)
    uint32_t space = 0;
    {
        const uint32_t hi = (c.cls >> 1) & 0x55555555u & ~c.cls;      // bit 2i set <=> class at position i is 2
        // compress the even bits to 16 bits
        uint32_t x = hi;
        x = (x | (x >> 1)) & 0x33333333u;
        x = (x | (x >> 2)) & 0x0F0F0F0Fu;
        x = (x | (x >> 4)) & 0x00FF00FFu;
        x = (x | (x >> 8)) & 0x0000FFFFu;
        space = x;
    }
    const uint32_t wstart = c.bnd & c.lead & ~space;
    uint32_t ev = (c.rows | wstart) & 0xFFFFu;
    int64_t nr = nr_hint;       // index of the first row that starts in this chunk, when the caller knows it
    row_first = row_last = nr < 0 ? 0 : nr;
    while (ev) {
#ifdef __CUDA_ARCH__
        const int i = __ffs(ev) - 1;
#else
        const int i = __builtin_ctz(ev);
#endif
        ev &= ev - 1u;
        const int64_t p = cs + i;
        if ((c.rows >> i) & 1u) {
            if (nr < 0) {
                nr = ak_row_lower_bound(X.off, X.r_lo, X.r_hi, p);
                row_first = nr;
            }
            while (nr <= X.n_rows && X.off[nr] == p) {
                if (nr > 0 && X.M->eos >= 0) ak_id_put(sink, X.M->eos);
                if (id_splits) id_splits[nr] = sink.cnt;
                if (nr < X.n_rows && X.M->bos >= 0) ak_id_put(sink, X.M->bos);
                ++nr;
            }
            row_last = nr;
        }
        if ((wstart >> i) & 1u) {
            const uint32_t k = (c.cls >> (2 * i)) & 3u;
            const int64_t e = akb_word_end(*X.T, X.text, cs, i, k, c.bnd, next_bnd, X.off, X.n_rows, X.r_lo, X.r_hi);
            akb_word(*X.M, *X.T, *X.C, X.text, p, e, k, sink, *X.pool, status);
        }
    }
}

// exact NFC check of the troubled code points of a chunk (cold): does NFC change the text?
AK_HD_NOINLINE bool akb_chunk_changes(const AkBLaneCtx& X, const AkBChunk& c, int64_t cs, int64_t limit, uint32_t& status) {
    uint32_t m = c.trb & 0xFFFFu;
    bool changed = false;
    int64_t checked_until = -1;
    while (m) {
#ifdef __CUDA_ARCH__
        const int i = __ffs(m) - 1;
#else
        const int i = __builtin_ctz(m);
#endif
        m &= m - 1u;
        const int64_t p = cs + i;
        if (p < checked_until) continue;
        const int64_t r = ak_row_lower_bound(X.off, X.r_lo, X.n_rows, p + 1);     // first row that starts after p
        const int64_t rs = X.off[r - 1], re = X.off[r];
        if (ak_segment_changes(*X.T, X.text, p, rs, re, limit, &checked_until, status)) changed = true;
    }
    return changed;
}
