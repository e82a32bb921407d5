// Unicode primitives shared by every kernel: UTF-8 cursor, property word lookup, canonical (de)composition.
// Everything is AK_HD so that the per-thread walkers can also be compiled by g++ for the CPU span tests
// (tests/csrc/host_harness.cpp) -- that build is a test aid, the product only ever runs the CUDA build.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define AK_HD __host__ __device__ __forceinline__
#define AK_HD_NOINLINE __host__ __device__ __noinline__
#else
#define AK_HD inline
#define AK_HD_NOINLINE inline
#endif

// property word layout: see tools/gen_tables.py
#define AK_GCB(w) ((w) & 15u)
#define AK_INCB(w) (((w) >> 4) & 3u)
#define AK_EXTPICT(w) (((w) >> 6) & 1u)
#define AK_TAG(w) (((w) >> 7) & 7u)
#define AK_ALLOW(w) (((w) >> 10) & 1u)
#define AK_QC(w) (((w) >> 11) & 3u)
#define AK_LATIN_LOWER(w) (((w) >> 13) & 1u)
#define AK_HFCLASS(w) (((w) >> 14) & 3u)
#define AK_CCC(w) (((w) >> 16) & 255u)
#define AK_FULL_LOWER(w) (((w) >> 24) & 1u)
#define AK_HAS_DECOMP(w) (((w) >> 25) & 1u)
#define AK_CASE_IGNORABLE(w) (((w) >> 26) & 1u)
#define AK_CASED(w) (((w) >> 27) & 1u)
#define AK_COMP_FIRST(w) (((w) >> 28) & 1u)
#define AK_BPE_SAFE(w) (((w) >> 29) & 1u)      // HF's NFKC (Unicode <= 12) acts on it like the NFC implemented here
// an atomic starter that can neither decompose nor compose with a following mark: a QC=Maybe mark right after it
// is left alone by NFC (e.g. Devanagari consonant + nukta, except U+0928 / U+0930 / U+0933)
#define AK_INERT_BASE(w) (((w) & ((255u << 16) | (3u << 11) | (1u << 25) | (1u << 28))) == 0u)
// a code point that starts an NFC segment: starter that neither decomposes nor composes backwards
#define AK_NFC_HEAD(w) (((w) & ((255u << 16) | (3u << 11))) == 0u)

enum { GCB_OTHER = 0, GCB_CR, GCB_LF, GCB_CONTROL, GCB_EXTEND, GCB_ZWJ, GCB_RI, GCB_PREPEND, GCB_SPACINGMARK,
       GCB_L, GCB_V, GCB_T, GCB_LV, GCB_LVT };
enum { INCB_NONE = 0, INCB_CONSONANT, INCB_LINKER, INCB_EXTEND };
enum { TAG_DEVANAGARI = 0, TAG_ROMAN, TAG_DIGIT, TAG_PUNCT, TAG_OTHER, TAG_NONE = 255 };

struct AkTables {
    const uint16_t* page_index;
    const uint32_t* leaves;
    const uint32_t* decomp_keys;
    const uint16_t* decomp_off;
    const uint32_t* decomp_data;
    const unsigned long long* pair_keys;
    const uint32_t* pair_vals;
    const uint32_t* ll_keys;
    const uint32_t* ll_vals;
    const uint32_t* fl_keys;
    const uint32_t* fl_vals;
    int n_decomp, n_pairs, n_ll, n_fl;
    // HF's NFKC (tools/gen_tables.py): compatibility decompositions as `tokenizers` applies them, and the code points its
    // older Unicode data does not know
    const uint32_t* kmap_keys;
    const uint16_t* kmap_off;
    const uint32_t* kmap_data;
    const uint32_t* hf_unknown;
    int n_kmap, n_hf_unknown;
};

AK_HD uint32_t ak_props(const AkTables& T, uint32_t cp) {
    if (cp >= 0x110000u) return 0u;
    return T.leaves[((uint32_t)T.page_index[cp >> 8] << 8) | (cp & 255u)];
}

AK_HD int ak_utf8_len(uint32_t cp) { return cp < 0x80u ? 1 : cp < 0x800u ? 2 : cp < 0x10000u ? 3 : 4; }

// decode the code point whose lead byte is at `pos`; never reads at or beyond `end`
AK_HD uint32_t ak_decode(const uint8_t* t, int64_t pos, int64_t end, int& len) {
    uint32_t b0 = t[pos];
    if (b0 < 0x80u) { len = 1; return b0; }
    // complete sequences straight-line (every walker decodes the same code points several times; the rolled loop below was
    // a tenth of the slow-lane kernel's instructions); the loop keeps the truncated tails
    if (b0 >= 0xE0u && b0 < 0xF0u && pos + 3 <= end) {
        len = 3;
        return ((b0 & 0x0Fu) << 12) | (((uint32_t)t[pos + 1] & 0x3Fu) << 6) | ((uint32_t)t[pos + 2] & 0x3Fu);
    }
    if (b0 >= 0xC0u && b0 < 0xE0u && pos + 2 <= end) {
        len = 2;
        return ((b0 & 0x1Fu) << 6) | ((uint32_t)t[pos + 1] & 0x3Fu);
    }
    if (b0 >= 0xF0u && pos + 4 <= end) {
        len = 4;
        return ((b0 & 0x07u) << 18) | (((uint32_t)t[pos + 1] & 0x3Fu) << 12) | (((uint32_t)t[pos + 2] & 0x3Fu) << 6) |
               ((uint32_t)t[pos + 3] & 0x3Fu);
    }
    int n = b0 >= 0xF0u ? 4 : b0 >= 0xE0u ? 3 : b0 >= 0xC0u ? 2 : 1;
    if (pos + n > end) n = (int)(end - pos);
    uint32_t cp = n == 1 ? b0 : (b0 & (0xFFu >> (n + 1)));
    for (int i = 1; i < n; ++i) cp = (cp << 6) | (t[pos + i] & 0x3Fu);
    len = n;
    return cp;
}

// start of the code point that ends right before `pos` (pos > lo)
AK_HD int64_t ak_prev_start(const uint8_t* t, int64_t pos, int64_t lo) {
    int64_t q = pos - 1;
    int k = 0;
    while (q > lo && k < 3 && (t[q] & 0xC0u) == 0x80u) { --q; ++k; }
    return q;
}

AK_HD int ak_encode(uint32_t cp, uint8_t* o) {
    if (cp < 0x80u) { o[0] = (uint8_t)cp; return 1; }
    if (cp < 0x800u) { o[0] = (uint8_t)(0xC0u | (cp >> 6)); o[1] = (uint8_t)(0x80u | (cp & 63u)); return 2; }
    if (cp < 0x10000u) {
        o[0] = (uint8_t)(0xE0u | (cp >> 12)); o[1] = (uint8_t)(0x80u | ((cp >> 6) & 63u)); o[2] = (uint8_t)(0x80u | (cp & 63u));
        return 3;
    }
    o[0] = (uint8_t)(0xF0u | (cp >> 18)); o[1] = (uint8_t)(0x80u | ((cp >> 12) & 63u));
    o[2] = (uint8_t)(0x80u | ((cp >> 6) & 63u)); o[3] = (uint8_t)(0x80u | (cp & 63u));
    return 4;
}

template <class K>
AK_HD int ak_bsearch(const K* keys, int n, K key) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        K v = keys[mid];
        if (v == key) return mid;
        if (v < key) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

// reference normalize.py:37-39 -- lower() of a code point whose Unicode name contains LATIN (first cp only;
// U+0130 additionally yields U+0307, handled by the caller)
AK_HD uint32_t ak_latin_lower(const AkTables& T, uint32_t cp) {
    if (cp >= 'A' && cp <= 'Z') return cp + 32u;
    int i = ak_bsearch<uint32_t>(T.ll_keys, T.n_ll, cp);
    return i < 0 ? cp : T.ll_vals[i];
}

AK_HD uint32_t ak_full_lower(const AkTables& T, uint32_t cp) {
    if (cp >= 'A' && cp <= 'Z') return cp + 32u;
    int i = ak_bsearch<uint32_t>(T.fl_keys, T.n_fl, cp);
    return i < 0 ? cp : T.fl_vals[i];
}

// ---- UAX #15 pieces (Unicode 15.0 data, probed from CPython's unicodedata) -------------------------------
#define AK_SB 0xAC00u
#define AK_LB 0x1100u
#define AK_VB 0x1161u
#define AK_TB 0x11A7u
#define AK_LC 19u
#define AK_VC 21u
#define AK_TC 28u
#define AK_NC (AK_VC * AK_TC)
#define AK_SC (AK_LC * AK_NC)

// full canonical decomposition of cp appended to buf[n..]; returns new n (caller guarantees room for 4)
AK_HD int ak_decompose(const AkTables& T, uint32_t cp, uint32_t props, uint32_t* buf, int n) {
    if (!AK_HAS_DECOMP(props)) { buf[n] = cp; return n + 1; }
    if (cp >= AK_SB && cp < AK_SB + AK_SC) {
        uint32_t s = cp - AK_SB;
        buf[n++] = AK_LB + s / AK_NC;
        buf[n++] = AK_VB + (s % AK_NC) / AK_TC;
        uint32_t t = s % AK_TC;
        if (t) buf[n++] = AK_TB + t;
        return n;
    }
    int i = ak_bsearch<uint32_t>(T.decomp_keys, T.n_decomp, cp);
    if (i < 0) { buf[n] = cp; return n + 1; }
    for (int k = T.decomp_off[i]; k < T.decomp_off[i + 1]; ++k) buf[n++] = T.decomp_data[k];
    return n;
}

AK_HD uint32_t ak_compose_pair(const AkTables& T, uint32_t a, uint32_t b) {
    if (a >= AK_LB && a < AK_LB + AK_LC && b >= AK_VB && b < AK_VB + AK_VC)
        return AK_SB + ((a - AK_LB) * AK_VC + (b - AK_VB)) * AK_TC;
    if (a >= AK_SB && a < AK_SB + AK_SC && (a - AK_SB) % AK_TC == 0 && b > AK_TB && b < AK_TB + AK_TC)
        return a + (b - AK_TB);
    unsigned long long key = ((unsigned long long)a << 21) | b;
    int i = ak_bsearch<unsigned long long>(T.pair_keys, T.n_pairs, key);
    return i < 0 ? 0u : T.pair_vals[i];
}

// NFC of an already fully-decomposed buffer: canonical ordering then canonical composition, in place.
AK_HD int ak_nfc_inplace(const AkTables& T, uint32_t* buf, int n) {
    // stable insertion sort of every maximal run of non-starters by ccc
    for (int i = 1; i < n; ++i) {
        uint32_t c = buf[i];
        uint32_t cc = AK_CCC(ak_props(T, c));
        if (cc == 0) continue;
        int j = i;
        while (j > 0) {
            uint32_t pc = AK_CCC(ak_props(T, buf[j - 1]));
            if (pc == 0 || pc <= cc) break;
            buf[j] = buf[j - 1];
            --j;
        }
        buf[j] = c;
    }
    int m = 0, starter = -1;
    int last_ccc = -1;
    for (int i = 0; i < n; ++i) {
        uint32_t c = buf[i];
        int cc = (int)AK_CCC(ak_props(T, c));
        if (starter >= 0 && (last_ccc == -1 || last_ccc < cc)) {
            uint32_t comp = ak_compose_pair(T, buf[starter], c);
            if (comp) { buf[starter] = comp; continue; }
        }
        if (cc == 0) { starter = m; last_ccc = -1; } else last_ccc = cc;
        buf[m++] = c;
    }
    return m;
}
