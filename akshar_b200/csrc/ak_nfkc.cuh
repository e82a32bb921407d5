// HF's NFKC (scripts/train_bpe.py:71 `normalizers.NFKC()`, tokenizers 0.22.2) for the rows the BPE encoder has to normalize
// itself: one row at a time, sequentially, in the row-fix kernel -- the fast path never sees a code point this changes.
//
// `tokenizers` carries older Unicode data than the NFC of normalize_text (CPython, Unicode 15): tools/gen_tables.py probes
// the normalizer code point by code point and leaves (1) kmap: its compatibility decomposition wherever that is not the
// canonical one, (2) hf_unknown: code points with a combining class or a decomposition here that HF passes through as
// inert starters.  NFKC = that decomposition, canonical ordering, canonical composition; a composite HF does not know
// stays decomposed.  tests/ check it against 3198 strings normalized by `tokenizers` itself (tests/golden).
#pragma once
#include "ak_text_core.cuh"

struct AkByteSink {
    uint8_t* out;           // nullptr: count only
    int64_t cnt, cap;
};
AK_HD void akk_put(AkByteSink& s, uint32_t cp) {
    uint8_t b[4];
    const int n = ak_encode(cp, b);
    if (s.out && s.cnt + n <= s.cap)
        for (int i = 0; i < n; ++i) s.out[s.cnt + i] = b[i];
    s.cnt += n;
}

AK_HD bool akk_unknown(const AkTables& T, uint32_t cp, uint32_t w) {
    return !AK_BPE_SAFE(w) && ak_bsearch<uint32_t>(T.hf_unknown, T.n_hf_unknown, cp) >= 0;
}

AK_HD void akk_flush(const AkTables& T, uint32_t* buf, int& n, AkByteSink& sink) {
    const int m = ak_nfc_inplace(T, buf, n);
    for (int i = 0; i < m; ++i) {
        const uint32_t c = buf[i];
        const uint32_t w = ak_props(T, c);
        if (AK_HAS_DECOMP(w) && akk_unknown(T, c, w)) {
            uint32_t d[4];
            const int k = ak_decompose(T, c, w, d, 0);
            for (int j = 0; j < k; ++j) akk_put(sink, d[j]);
        } else akk_put(sink, c);
    }
    n = 0;
}

// NFKC of text [s, e) (whole code points) appended to sink
AK_HD_NOINLINE void akk_nfkc(const AkTables& T, const uint8_t* t, int64_t s, int64_t e, AkByteSink& sink, uint32_t& status) {
    uint32_t buf[AK_MAXSEG];
    int n = 0;
    int64_t q = s;
    while (q < e) {
        int len;
        const uint32_t cp = ak_decode(t, q, e, len);
        q += len;
        const uint32_t w = ak_props(T, cp);
        if (AK_BPE_SAFE(w) && AK_NFC_HEAD(w) && !AK_HAS_DECOMP(w)) {       // nearly every code point: nothing to look up
            if (n) akk_flush(T, buf, n, sink);
            buf[n++] = cp;
            continue;
        }
        if (akk_unknown(T, cp, w)) {
            if (n) akk_flush(T, buf, n, sink);
            akk_put(sink, cp);
            continue;
        }
        uint32_t d[20];
        int k = 0;
        const int ki = AK_BPE_SAFE(w) ? -1 : ak_bsearch<uint32_t>(T.kmap_keys, T.n_kmap, cp);
        if (ki >= 0) {
            for (int j = T.kmap_off[ki]; j < T.kmap_off[ki + 1] && k < 20; ++j) d[k++] = T.kmap_data[j];
        } else k = ak_decompose(T, cp, w, d, 0);
        if (k > 0 && AK_NFC_HEAD(ak_props(T, d[0])) && n) akk_flush(T, buf, n, sink);
        if (n + k > AK_MAXSEG) {
            // a run of marks longer than the segment buffer: ordering / composition across the cut is not attempted
            status |= AK_ST_NFC_SEGMENT;
            akk_flush(T, buf, n, sink);
        }
        for (int j = 0; j < k; ++j) buf[n++] = d[j];
    }
    if (n) akk_flush(T, buf, n, sink);
}
