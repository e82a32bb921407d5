// Word cache of the subword encoders: pre-tokenized word (its bytes) -> token ids.
//
// HF tokenizers keeps such a cache inside its BPE model (reference tokenizer.py:96-98 -> `Tokenizer.from_file`, the BPE
// `cache`); here it is an open-addressing table in global memory (L2 resident, the hot entries L1 resident) shared by the
// BPE encoder (word = one `\w+` / `[^\w\s]+` run, value = its merged ids) and the Unigram encoder (word = one run of
// non-space bytes, value = the Viterbi segmentation of U+2581 + word, plus the two numbers that say when that
// segmentation is independent of the score accumulated before the word -- see ak_tok.cuh).
//
// Entry = 16 x u64 = 128 bytes = four 32-byte sectors, laid out so that ONE sector answers the common case (a word of at
// most 16 bytes with at most 2 ids):
//   w0  tag    [63:48] / [47:32] two bf16 numbers (Unigram: margin / length rounded down, largest |partial score| of the
//              word's lattice rounded up; see ak_tok.cuh), [31:16] hash, [15:8] byte length, [7:3] id count, bit 1 busy,
//              bit 0 ready
//   w1  w2     key bytes 0..15
//   w3         ids 0, 1
//   w4..w8     key bytes 16..55
//   w9..w14    ids 2..13
//   w15        unused
// Entries only ever go from empty to ready (lock-free: CAS on the tag, payload, fence, tag), so look-ups are ordinary
// cacheable loads: a stale sector can only show "still empty", which is a miss that is recomputed exactly.
// Everything is AK_HD: the image is built on the host at model load with the same functions the kernels use.
#pragma once
#include "ak_unicode.cuh"

#define AKC_ENTRY 16
#define AKC_MAXLEN 56
#define AKC_MAXTOK 14
#define AKC_PROBES 16
#define AKC_READY 1ull
#define AKC_BUSY 2ull
#define AKC_NTOK(tag) ((int)(((tag) >> 3) & 31ull))
#define AKC_K1_SALT 0x9E3779B97F4A7C15ull

struct AkWordCache {
    unsigned long long* e;          // (1 << bits) entries of AKC_ENTRY x u64
    uint32_t bits;
    unsigned long long* inserted;   // optional counter of the entries added since the image was last restored
};

AK_HD unsigned long long akc_ld(const unsigned long long* p) {
#ifdef __CUDA_ARCH__
    unsigned long long v;
    asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#else
    return *p;
#endif
}
AK_HD unsigned long long akc_ld_fresh(const unsigned long long* p) {
#ifdef __CUDA_ARCH__
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#else
    return *p;
#endif
}
AK_HD void akc_st(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
    *p = v;
#endif
}

// key word j (bytes [8j, 8j + 8) of the word [s, s + len), zero padded, little endian).  Device: two 8-byte aligned
// loads + a funnel shift (every CUDA allocation reaches the next 8-byte boundary after its last byte).
AK_HD unsigned long long akc_key_word(const uint8_t* t, int64_t s, uint32_t len, uint32_t j) {
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

// the first four key words (32 bytes) without a branch: nine aligned 32-bit loads (clamped to the last word of the text,
// whose bytes are masked off anyway), eight funnel shifts, byte masks from the length
AK_HD void akc_key0123(const uint8_t* t, int64_t s, uint32_t len, int64_t te, unsigned long long* k) {
#ifdef __CUDA_ARCH__
    const uintptr_t a = (uintptr_t)(t + s);
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t* last = (const uint32_t*)(((uintptr_t)(t + te) - 1u) & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3u) * 8u;
    uint32_t x[9];
    if (w + 8 <= last) {                                            // everywhere but in the last 36 bytes of the text
#pragma unroll
        for (int i = 0; i < 9; ++i) x[i] = __ldg(w + i);
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) x[i] = __ldg(w + i <= last ? w + i : last);
    }
    uint32_t f[8];
    const int len8 = 8 * (int)len;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // bits of the word in this 32-bit piece: 0 .. 32 and beyond.  PTX shl clamps the amount at 32 (-> 0, mask of ones)
        const int bits = max(len8 - 32 * i, 0);
        uint32_t m;
        asm("shl.b32 %0, %1, %2;" : "=r"(m) : "r"(1u), "r"((uint32_t)bits));
        f[i] = __funnelshift_r(x[i], x[i + 1], sh) & (m - 1u);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = ((unsigned long long)f[2 * j + 1] << 32) | f[2 * j];
#else
    (void)te;
    for (uint32_t j = 0; j < 4; ++j) k[j] = len > 8u * j ? akc_key_word(t, s, len, j) : 0ull;
#endif
}
// (host image builder)
AK_HD void akc_key01(const uint8_t* t, int64_t s, uint32_t len, int64_t te, unsigned long long& k0, unsigned long long& k1) {
    unsigned long long k[4];
    akc_key0123(t, s, len, te, k);
    k0 = k[0];
    k1 = k[1];
}

AK_HD unsigned long long akc_mix(unsigned long long h, unsigned long long k) {
    h = (h ^ k) * 0xBF58476D1CE4E5B9ull;
    return (h << 27) | (h >> 37);
}
AK_HD unsigned long long akc_fin(unsigned long long h) {
    h ^= h >> 31;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    return h;
}
AK_HD unsigned long long akc_want(unsigned long long h, uint32_t len) { return ((h >> 48) << 16) | ((unsigned long long)len << 8) | AKC_READY; }
#define AKC_MATCH_MASK 0xFFFFFF07ull        // hash, length, state: what a look-up compares (not the id count, not the aux numbers)
// the two bf16 numbers of the tag (Unigram): a conservative float pair
AK_HD unsigned long long akc_aux(float ratio_down, float wmag_up) {
    union { float f; uint32_t u; } a, b;
    a.f = ratio_down;
    b.f = wmag_up;
    const uint32_t r = a.u >> 16;                               // truncation = towards zero (ratio >= 0)
    uint32_t w = (b.u + 0xFFFFu) >> 16;                         // round up
    if (w > 0x7F7Fu) w = 0x7F7Fu;                               // stay finite
    return ((unsigned long long)r << 48) | ((unsigned long long)w << 32);
}
AK_HD float akc_ratio(unsigned long long tag) {
    union { float f; uint32_t u; } a;
    a.u = (uint32_t)(tag >> 48) << 16;
    return a.f;
}
AK_HD float akc_wmag(unsigned long long tag) {
    union { float f; uint32_t u; } a;
    a.u = (uint32_t)((tag >> 32) & 0xFFFFu) << 16;
    return a.f;
}

struct AkcHit {
    long long slot;                 // entry index, -1 = miss
    long long free_slot;            // on a miss: an empty slot on the probe path, or -1
    unsigned long long h, want;
    unsigned long long tag, ids01;  // on a hit: the tag (id count) and the first two ids
};

// hash + probe; len <= AKC_MAXLEN.  k[0..3] = the first four key words (akc_key0123).
AK_HD void akc_lookup4(const AkWordCache& C, const uint8_t* t, int64_t s, uint32_t len, const unsigned long long* k, AkcHit& r) {
    const uint32_t nw = (len + 7u) >> 3;
    unsigned long long h = 0x9E3779B97F4A7C15ull + len;
    h = akc_mix(h, k[0]);
    if (nw > 1u) h = akc_mix(h, k[1]);
    if (nw > 2u) h = akc_mix(h, k[2]);
    if (nw > 3u) h = akc_mix(h, k[3]);
    if (nw > 4u) {
#pragma unroll 1
        for (uint32_t j = 4; j < nw; ++j) h = akc_mix(h, akc_key_word(t, s, len, j));
    }
    h = akc_fin(h);
    const unsigned long long want = akc_want(h, len);
    r.h = h;
    r.want = want;
    r.slot = -1;
    r.free_slot = -1;
    r.tag = 0ull;
    r.ids01 = 0ull;
    const unsigned long long mask = (1ull << C.bits) - 1ull;
#pragma unroll 1
    for (int p = 0; p < AKC_PROBES; ++p) {
        const unsigned long long slot = (h + (unsigned long long)p) & mask;
        const unsigned long long* e = C.e + slot * AKC_ENTRY;
#ifdef __CUDA_ARCH__
        // the entry's first sector as two 16-byte loads
        ulonglong2 a, b;
        asm volatile("ld.global.ca.v2.u64 {%0, %1}, [%2];" : "=l"(a.x), "=l"(a.y) : "l"(e) : "memory");
        asm volatile("ld.global.ca.v2.u64 {%0, %1}, [%2];" : "=l"(b.x), "=l"(b.y) : "l"(e + 2) : "memory");
        const unsigned long long tag = a.x, e0 = a.y, e1 = b.x, i01 = b.y;
#else
        const unsigned long long tag = e[0], e0 = e[1], e1 = e[2], i01 = e[3];
#endif
        if (tag == 0ull) { r.free_slot = (long long)slot; return; }
        // the two loads are independent requests: (k1 ^ AKC_K1_SALT) is stored last but the tag, behind a fence, so a
        // second half that is older than the first one never matches (an all-zero half cannot: the salt is not UTF-8)
        if ((tag & AKC_MATCH_MASK) != want || e0 != k[0] || e1 != (k[1] ^ AKC_K1_SALT)) continue;
        bool same = true;
        if (nw > 2u) {
            // key bytes 16..31: half of the entry's second sector
#ifdef __CUDA_ARCH__
            ulonglong2 c;
            asm volatile("ld.global.ca.v2.u64 {%0, %1}, [%2];" : "=l"(c.x), "=l"(c.y) : "l"(e + 4) : "memory");
            same = c.x == k[2] && (nw <= 3u || c.y == k[3]);
#else
            same = e[4] == k[2] && (nw <= 3u || e[5] == k[3]);
#endif
            if (same && nw > 4u) {
#pragma unroll 1
                for (uint32_t j = 4; j < nw; ++j)
                    if (akc_ld(e + 2 + j) != akc_key_word(t, s, len, j)) { same = false; break; }
            }
        }
        if (same) {
            r.slot = (long long)slot;
            r.tag = tag;
            r.ids01 = i01;
            return;
        }
    }
}
AK_HD void akc_lookup(const AkWordCache& C, const uint8_t* t, int64_t s, uint32_t len, unsigned long long k0, unsigned long long k1,
                      AkcHit& r) {
    unsigned long long k[4] = {k0, k1, len > 16u ? akc_key_word(t, s, len, 2) : 0ull, len > 24u ? akc_key_word(t, s, len, 3) : 0ull};
    akc_lookup4(C, t, s, len, k, r);
}

// id i (>= 2) of entry e
AK_HD uint32_t akc_id(const unsigned long long* e, int i) {
    const unsigned long long v = akc_ld(e + 8 + (i >> 1));
    return (i & 1) ? (uint32_t)(v >> 32) : (uint32_t)v;
}

AK_HD void akc_insert(const AkWordCache& C, long long slot, unsigned long long want, const uint8_t* t, int64_t s, uint32_t len,
                      const int32_t* ids, int n, unsigned long long aux) {
    unsigned long long* e = C.e + (unsigned long long)slot * AKC_ENTRY;
#ifdef __CUDA_ARCH__
    if (atomicCAS(e, 0ull, AKC_BUSY) != 0ull) return;
    if (C.inserted) atomicAdd(C.inserted, 1ull);
#else
    if (*e != 0ull) return;
    *e = AKC_BUSY;
    if (C.inserted) *C.inserted += 1ull;
#endif
    const uint32_t nw = (len + 7u) >> 3;
    akc_st(e + 1, akc_key_word(t, s, len, 0));
    for (uint32_t j = 2; j < nw; ++j) akc_st(e + 2 + j, akc_key_word(t, s, len, j));
    for (int i = 0; i < n; i += 2) {
        unsigned long long v = (uint32_t)ids[i];
        if (i + 1 < n) v |= (unsigned long long)(uint32_t)ids[i + 1] << 32;
        akc_st(e + (i == 0 ? 3 : 8 + (i >> 1)), v);
    }
    if (n == 0) akc_st(e + 3, 0ull);
#ifdef __CUDA_ARCH__
    __threadfence();
#endif
    akc_st(e + 2, (nw > 1u ? akc_key_word(t, s, len, 1) : 0ull) ^ AKC_K1_SALT);
#ifdef __CUDA_ARCH__
    __threadfence();
#endif
    akc_st(e, want | ((unsigned long long)n << 3) | aux);
}
