// Host side of the on-device decode: the piece tables (ak_decode.cuh) built at model load.  Included by ak_kernels.cu and by
// the CPU test harness.
#pragma once
#include <string>
#include <vector>

#include "ak_decode.cuh"
#include "ak_models.h"

struct AkDecHost {
    std::vector<uint32_t> off;
    std::vector<uint8_t> bytes, flags;
    int32_t strict = 0;
    AkDecTable view() const {
        AkDecTable t;
        t.off = off.data();
        t.bytes = bytes.data();
        t.flags = flags.data();
        t.size = (int32_t)flags.size();
        t.strict = strict;
        return t;
    }
    void put(const std::string& s, uint8_t f) {
        bytes.insert(bytes.end(), s.begin(), s.end());
        off.push_back((uint32_t)bytes.size());
        flags.push_back(f);
    }
};

inline std::string akd_replace(std::string s, const std::string& from, const std::string& to) {
    for (size_t p = 0; (p = s.find(from, p)) != std::string::npos; p += to.size()) s.replace(p, from.size(), to);
    return s;
}
inline bool akd_all_space(const std::string& s, size_t from) {
    const uint8_t* b = (const uint8_t*)s.data();
    for (int64_t q = (int64_t)from; q < (int64_t)s.size();) {
        int len;
        if (!akd_isspace(ak_decode(b, q, (int64_t)s.size(), len))) return false;
        q += len;
    }
    return true;
}

// HF tokenizers JSON (scripts/train_bpe.py:68-98; `decoder: null`).  decode: the tokens of the ids that are not special,
// joined by one space (tokenizer.py:219-220); detokenize (tokenizer.py:240-244): ' '.join, ' ##' removed, U+0120 -> ' ', strip
inline AkDecHost ak_build_bpe_decode(const AkBpeHost& h, int form) {
    AkDecHost t;
    t.off.push_back(0);
    for (size_t id = 0; id < h.id_to_token.size(); ++id) {
        const std::string& tok = h.id_to_token[id];
        if (form == AKD_FORM_DECODE) {
            if (h.is_special[id] || tok.empty()) t.put("", AKD_SKIP);        // ids the vocabulary does not use decode to nothing
            else t.put(" " + tok, AKD_LEADSP);
            continue;
        }
        const std::string body = akd_replace(tok, "\xC4\xA0", " ");
        if (tok.compare(0, 2, "##") == 0) {
            uint8_t f = AKD_HH;
            if (akd_all_space(body, 2)) f |= AKD_WS_REST;                   // first in its row it keeps the "##"
            t.put(body, f);
        } else {
            uint8_t f = AKD_LEADSP;
            if (akd_all_space(body, 0)) f |= AKD_WS_FIRST | AKD_WS_REST;
            t.put(" " + body, f);
        }
    }
    return t;
}

// SentencePiece ModelProto (scripts/train_spm.py:80-108).  decode = DecodeIds (tokenizer.py:217-218): control pieces vanish,
// <unk> is its surface " ⁇ ", byte pieces are reassembled, U+2581 -> ' ', one leading U+2581 consumed while the text is
// empty; detokenize (tokenizer.py:236-239): the piece strings concatenated, U+2581 -> ' ', strip
inline AkDecHost ak_build_spm_decode(const AkUniHost& h, int form) {
    AkDecHost t;
    t.off.push_back(0);
    t.strict = 1;
    const std::string us = "\xE2\x96\x81";
    for (size_t id = 0; id < h.piece.size(); ++id) {
        const std::string& p = h.piece[id];
        const int ty = h.type[id];
        if (form == AKD_FORM_DETOK) {
            const std::string body = akd_replace(p, us, " ");
            t.put(body, akd_all_space(body, 0) ? (uint8_t)(AKD_WS_FIRST | AKD_WS_REST) : (uint8_t)0);
            continue;
        }
        if (ty == 3) { t.put("", AKD_SKIP); continue; }                                  // CONTROL
        if (ty == 2) { t.put(" \xE2\x81\x87 ", 0); continue; }                           // UNKNOWN: unk_surface
        if (ty == 6 && p.size() == 6) {                                                  // BYTE "<0xNN>"
            t.put(std::string(1, (char)strtol(p.substr(3, 2).c_str(), nullptr, 16)), AKD_BYTE);
            continue;
        }
        uint8_t f = 0;
        if (p.compare(0, 3, us) == 0) f |= AKD_LEADSP;
        if (p.empty() || p == us) f |= AKD_BOSEMPTY;
        t.put(akd_replace(p, us, " "), f);
    }
    return t;
}
