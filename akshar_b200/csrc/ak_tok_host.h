// Host side of the event-stream encoders: the word-cache IMAGES built at model load (ak_wordcache.cuh), with the same
// AK_HD functions the kernels run.  Included by ak_kernels.cu and by the CPU test harness.
#pragma once
#include <string>
#include <vector>

#include "ak_models.h"
#include "ak_tok.cuh"

#define AKC_BITS 19        // 2^19 entries x 128 bytes = 64 MiB per model: room for the vocabulary and 128 k learned words

// host views of the uploaded model tables
inline AkBpeDev ak_bpe_host_view(const AkBpeHost& h) {
    AkBpeDev m{};
    m.cp_direct = h.cp_direct.data();
    m.cp_keys = h.cp_keys.data();
    m.cp_ids = h.cp_ids.data();
    m.n_cp = (int)h.cp_keys.size();
    m.mkeys = h.mkeys.data();
    m.mvals = h.mvals.data();
    m.mbits = h.mbits;
    m.bos = h.bos;
    m.eos = h.eos;
    m.sp_bytes = h.sp_bytes.data();
    m.sp_off = h.sp_off.data();
    m.sp_ids = h.sp_ids.data();
    m.n_sp = (int)h.sp_ids.size();
    return m;
}
inline AkUniDev ak_uni_host_view(const AkUniHost& h) {
    AkUniDev u{};
    u.tkeys = h.tkeys.data();
    u.tvals = h.tvals.data();
    u.tkv = nullptr;
    u.tbits = h.tbits;
    u.score = h.score.data();
    u.usable = h.usable.data();
    u.byte_id = h.byte_id;
    u.unk_id = h.unk_id;
    u.unk_score = h.unk_score;
    u.flags = h.flags;
    return u;
}

inline void ak_image_put(AkWordCache& hc, const uint8_t* tb, uint32_t n, const int32_t* ids, int cnt, unsigned long long aux) {
    unsigned long long k[4];
    akc_key0123(tb, 0, n, (int64_t)n, k);
    AkcHit h;
    akc_lookup4(hc, tb, 0, n, k, h);
    if (h.slot < 0 && h.free_slot >= 0) akc_insert(hc, h.free_slot, h.want, tb, 0, n, ids, cnt, aux);
}

// BPE: every vocabulary string that is exactly one pre-tokenizer word, encoded by the merge loop
inline std::vector<unsigned long long> ak_build_bpe_image(const AkBpeHost& h, const AkTables& ht, uint32_t bits) {
    const AkBpeDev hm = ak_bpe_host_view(h);
    std::vector<unsigned long long> img((size_t)AKC_ENTRY << bits, 0ull);
    AkWordCache hc;
    hc.e = img.data();
    hc.bits = bits;
    hc.inserted = nullptr;
    std::vector<int32_t> poolbuf(4096);
    unsigned long long used = 0;
    AkPool hp;
    hp.base = poolbuf.data();
    hp.used = &used;
    hp.cap = poolbuf.size();
    for (size_t id = 0; id < h.id_to_token.size(); ++id) {
        const std::string& tok = h.id_to_token[id];
        if (tok.empty() || tok.size() > AKC_MAXLEN || h.is_special[id]) continue;
        const uint8_t* tb = (const uint8_t*)tok.data();
        const int64_t n = (int64_t)tok.size();
        uint32_t k = 3;
        bool one_word = true;
        for (int64_t q = 0; q < n;) {
            int len;
            const uint32_t cp = ak_decode(tb, q, n, len);
            const uint32_t kk = AK_HFCLASS(ak_props(ht, cp));
            if (kk == 2u || (k != 3u && kk != k)) { one_word = false; break; }
            k = kk;
            q += len;
        }
        if (!one_word || k == 3u) continue;
        used = 0;
        uint32_t st = 0;
        AkIdSink sink;
        int32_t out_ids[AKC_MAXTOK + 1];
        sink.buf = out_ids;
        sink.cap = AKC_MAXTOK + 1;
        sink.stride = 1;
        sink.cnt = 0;
        sink.direct = false;
        sink.gout = nullptr;
        sink.gbase = 0;
        sink.gcap = 0;
        ak_bpe_word(hm, ht, tb, 0, n, k, sink, hp, st);
        if (st || sink.cnt > AKC_MAXTOK) continue;
        ak_image_put(hc, tb, (uint32_t)n, out_ids, sink.cnt, 0ull);
    }
    return img;
}

// Unigram: the word-wise path needs the model shape scripts/train_spm.py:80-108 produces: dummy prefix, extra white space
// removed, spaces escaped, and no piece with U+2581 anywhere but in front (split_by_whitespace at training time)
inline bool ak_uni_wordwise(const AkUniHost& h) {
    if ((h.flags & 7) != 7) return false;
    for (size_t i = 0; i < h.piece.size(); ++i) {
        if (h.type[i] != 1 && h.type[i] != 4 && h.type[i] != 5) continue;
        const std::string& p = h.piece[i];
        if (p.find("\xE2\x96\x81", 1) != std::string::npos) return false;
        if (p.find(' ') != std::string::npos) return false;
    }
    return true;
}
// every piece that starts with U+2581 is a word of its own: solve its lattice once
inline std::vector<unsigned long long> ak_build_uni_image(const AkUniHost& h, uint32_t bits) {
    const AkUniDev hu = ak_uni_host_view(h);
    std::vector<unsigned long long> img((size_t)AKC_ENTRY << bits, 0ull);
    AkWordCache hc;
    hc.e = img.data();
    hc.bits = bits;
    hc.inserted = nullptr;
    AkUniWord W;
    for (size_t i = 0; i < h.piece.size(); ++i) {
        if (h.type[i] != 1) continue;
        const std::string& p = h.piece[i];
        if (p.size() <= 3 || p.compare(0, 3, "\xE2\x96\x81") != 0 || p.size() - 3 > AKC_MAXLEN) continue;
        const uint8_t* tb = (const uint8_t*)p.data() + 3;
        const uint32_t n = (uint32_t)(p.size() - 3);
        aku_word_lattice(hu, tb, 0, n, W);
        if (!W.ok || W.n_ids > AKC_MAXTOK) continue;
        ak_image_put(hc, tb, n, W.ids, W.n_ids, akc_aux(W.ratio, W.wmag));
    }
    return img;
}
