// ids -> text on the device (reference tokenizer.py:195-246): `decode` (SentencePiece DecodeIds / HF Tokenizer.decode with
// `decoder: null`, tokenizer.py:217-220) and `detokenize` applied to the pieces of the ids (tokenizer.py:236-246).
//
// Both are a gather from a per-model PIECE TABLE built at load (ak_decode_host.h) whose entries are already in output
// form -- U+2581 / U+0120 replaced by a space, the BPE joiner in front, <unk> as its surface, a byte piece as its byte --
// plus four position-dependent rules, all decided per id from a one-byte MARK that a row-walk kernel leaves:
//   FIRST    the entry's leading joiner / dummy-prefix space is dropped: the first text of a row (HF joins with ' ';
//            SentencePiece consumes one leading U+2581 of every piece for as long as the decoded text is empty)
//   DROP / LSTRIP / RSTRIP   str.strip() of `detokenize`: pieces that are white space only at either end of the row
//            vanish, the first / last piece with anything else loses its leading / trailing white space
//   ROWSTART a row begins at this id (runs of byte pieces do not cross rows)
// Runs of SentencePiece byte pieces are reassembled as UTF-8: a byte that is part of a well-formed sequence inside its
// run comes out as it is, any other byte as U+FFFD (observed from sentencepiece 0.2.1; tests/golden decode_fuzz).
// AK_HD throughout: tests/csrc/host_harness.cpp runs the same code on the CPU.
#pragma once
#include "ak_unicode.cuh"
#include "ak_text_core.cuh"
#include "ak_wordtok.cuh"

#define AKD_FORM_DECODE 0
#define AKD_FORM_DETOK 1

// table flags
#define AKD_SKIP 1u             // no output (control piece / special token)
#define AKD_BYTE 2u             // SentencePiece byte piece: the entry is the byte
#define AKD_LEADSP 4u           // the entry's first byte is a joiner / dummy-prefix space
#define AKD_BOSEMPTY 8u         // yields nothing at the start of the text ('▁', ''): the text is still empty after it
#define AKD_HH 32u              // BPE detokenize: the piece starts with "##" (joined without the joiner and without the "##")
#define AKD_WS_FIRST 64u        // detokenize: white space only when it is the row's first piece
#define AKD_WS_REST 128u        // ... when it is not

// marks
#define AKD_M_FIRST 1u
#define AKD_M_DROP 2u
#define AKD_M_LSTRIP 4u
#define AKD_M_RSTRIP 8u
#define AKD_M_ROWSTART 16u

struct AkDecTable {
    const uint32_t* off;        // [size + 1]
    const uint8_t* bytes;
    const uint8_t* flags;       // [size]
    int32_t size;
    int32_t strict;             // ids outside [0, size) are an error (SentencePiece) instead of skipped (HF)
};

AK_HD bool akd_isspace(uint32_t cp) {
    return (cp >= 0x09u && cp <= 0x0Du) || (cp >= 0x1Cu && cp <= 0x20u) || (cp >= 0x80u && akw_isspace_wide(cp));
}

AK_HD uint32_t akd_flags(const AkDecTable& D, int64_t id) {
    return (id < 0 || id >= D.size) ? AKD_SKIP : D.flags[id];
}

// the row walk: marks of row [lo, hi) of ids.  One thread per row; rows are short, and the walks stop at the first piece
// that has text (decode) / that is not white space (detokenize).
template <class IdT>
AK_HD void akd_mark_row(const AkDecTable& D, int form, const IdT* ids, int64_t lo, int64_t hi, uint8_t* mark, uint32_t& st) {
    if (lo >= hi) return;
    mark[lo] |= AKD_M_ROWSTART;
    for (int64_t i = lo; i < hi; ++i) {
        const int64_t id = (int64_t)ids[i];
        if (id < 0 || id >= D.size) continue;
        const uint32_t f = D.flags[id];
        if (f & AKD_SKIP) continue;
        if (f & AKD_BYTE) break;
        if (f & (AKD_LEADSP | AKD_HH)) mark[i] |= AKD_M_FIRST;
        if (!(f & AKD_BOSEMPTY)) break;
    }
    if (form != AKD_FORM_DETOK) return;
    int64_t first_kept = hi;
    for (int64_t i = lo; i < hi; ++i) {
        const uint32_t f = akd_flags(D, (int64_t)ids[i]);
        if (f & (i == lo ? AKD_WS_FIRST : AKD_WS_REST)) { mark[i] |= AKD_M_DROP; continue; }
        mark[i] |= AKD_M_LSTRIP;
        first_kept = i;
        break;
    }
    for (int64_t i = hi - 1; i >= first_kept; --i) {
        const uint32_t f = akd_flags(D, (int64_t)ids[i]);
        if (i > first_kept && (f & (i == lo ? AKD_WS_FIRST : AKD_WS_REST))) { mark[i] |= AKD_M_DROP; continue; }
        mark[i] |= AKD_M_RSTRIP;
        break;
    }
}

// the part [start, end) of id's entry that is written
AK_HD void akd_range(const AkDecTable& D, int64_t id, uint32_t f, uint32_t m, uint32_t& start, uint32_t& end) {
    start = end = 0;
    if ((f & AKD_SKIP) || (m & AKD_M_DROP)) return;
    const uint32_t o = D.off[id];
    const uint32_t n = D.off[id + 1] - o;
    uint32_t s = 0, e = n;
    if ((f & AKD_HH) && !(m & AKD_M_FIRST)) s = 2;
    if ((f & AKD_LEADSP) && (m & AKD_M_FIRST)) s = 1;
    if (m & (AKD_M_LSTRIP | AKD_M_RSTRIP)) {
        const uint8_t* b = D.bytes + o;
        if (m & AKD_M_LSTRIP) {
            while (s < e) {
                int len;
                const uint32_t cp = ak_decode(b, s, e, len);
                if (!akd_isspace(cp)) break;
                s += (uint32_t)len;
            }
        }
        if (m & AKD_M_RSTRIP) {
            while (e > s) {
                uint32_t q = e - 1;
                while (q > s && (b[q] & 0xC0u) == 0x80u) --q;
                int len;
                const uint32_t cp = ak_decode(b, q, e, len);
                if (!akd_isspace(cp)) break;
                e = q;
            }
        }
    }
    start = s;
    end = e > s ? e : s;
}

// ---- runs of byte pieces -------------------------------------------------------------------------------------------
// length of the well-formed UTF-8 sequence that starts with lead byte v given the byte after it (0 = none)
AK_HD int akd_seq_len(uint32_t v, uint32_t b1) {
    if (v >= 0xC2u && v <= 0xDFu) return 2;
    if (v >= 0xE0u && v <= 0xEFu) {
        const uint32_t lo = v == 0xE0u ? 0xA0u : 0x80u, hi = v == 0xEDu ? 0x9Fu : 0xBFu;
        return (b1 >= lo && b1 <= hi) ? 3 : 0;
    }
    if (v >= 0xF0u && v <= 0xF4u) {
        const uint32_t lo = v == 0xF0u ? 0x90u : 0x80u, hi = v == 0xF4u ? 0x8Fu : 0xBFu;
        return (b1 >= lo && b1 <= hi) ? 4 : 0;
    }
    return 0;
}

// byte of id i if it is a byte piece of the same row as the walk that asks, else -1
template <class IdT>
AK_HD int akd_byte_at(const AkDecTable& D, const IdT* ids, int64_t i) {
    const int64_t id = (int64_t)ids[i];
    if (id < 0 || id >= D.size || !(D.flags[id] & AKD_BYTE)) return -1;
    return (int)D.bytes[D.off[id]];
}

// does a well-formed sequence start at byte piece j (lead byte v)?  The run ends at n_ids, at the next row or at the first
// id that is not a byte piece.
template <class IdT>
AK_HD int akd_valid_from(const AkDecTable& D, const IdT* ids, const uint8_t* mark, int64_t n_ids, int64_t j, uint32_t v) {
    if (v < 0xC2u || v > 0xF4u) return 0;
    int b[3] = {-1, -1, -1};
    for (int k = 1; k <= 3; ++k) {
        if (j + k >= n_ids || (mark[j + k] & AKD_M_ROWSTART)) break;
        b[k - 1] = akd_byte_at(D, ids, j + k);
        if (b[k - 1] < 0) break;
    }
    if (b[0] < 0) return 0;
    const int n = akd_seq_len(v, (uint32_t)b[0]);
    if (n == 0) return 0;
    for (int k = 1; k < n; ++k)
        if (b[k - 1] < 0 || (b[k - 1] & 0xC0) != 0x80) return 0;
    return n;
}

// what byte piece i (byte v) contributes: true = the byte itself, false = U+FFFD
template <class IdT>
AK_HD bool akd_byte_kept(const AkDecTable& D, const IdT* ids, const uint8_t* mark, int64_t n_ids, int64_t i, uint32_t v) {
    if (v < 0x80u) return true;
    if ((v & 0xC0u) != 0x80u) return akd_valid_from(D, ids, mark, n_ids, i, v) != 0;
    // a continuation byte: kept when the nearest lead before it (within its run) starts a well-formed sequence that reaches it
    for (int k = 1; k <= 3; ++k) {
        const int64_t j = i - k;
        if (j < 0 || (mark[j + 1] & AKD_M_ROWSTART)) return false;
        const int b = akd_byte_at(D, ids, j);
        if (b < 0 || b < 0x80) return false;
        if ((b & 0xC0) != 0x80) return akd_valid_from(D, ids, mark, n_ids, j, (uint32_t)b) > k;
    }
    return false;
}

// output length of id i; `bytes` (optional) receives them
template <class IdT>
AK_HD uint32_t akd_emit(const AkDecTable& D, const IdT* ids, const uint8_t* mark, int64_t n_ids, int64_t i, uint8_t* out, uint32_t& st) {
    const int64_t id = (int64_t)ids[i];
    if (id < 0 || id >= D.size) {
        if (D.strict) st |= AK_ST_BAD_ID;
        return 0;
    }
    const uint32_t f = D.flags[id];
    const uint32_t m = mark[i];
    if (f & AKD_BYTE) {
        const uint32_t v = D.bytes[D.off[id]];
        if (akd_byte_kept(D, ids, mark, n_ids, i, v)) {
            if (out) out[0] = (uint8_t)v;
            return 1;
        }
        if (out) { out[0] = 0xEF; out[1] = 0xBF; out[2] = 0xBD; }
        return 3;
    }
    uint32_t s, e;
    akd_range(D, id, f, m, s, e);
    if (out) {
        const uint8_t* b = D.bytes + D.off[id];
        for (uint32_t q = s; q < e; ++q) out[q - s] = b[q];
    }
    return e - s;
}
