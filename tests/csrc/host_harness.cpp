// CPU build of the span walkers (akshar_b200/csrc/ak_text_core.cuh) -- TEST AID ONLY.
// It lets tests/test_span_walkers.py check, without a GPU, that cutting the buffer into arbitrary spans never
// changes the result (the property the CUDA kernels rely on).  Not linked into libakshar_b200.so.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <math.h>
#include "../../akshar_b200/csrc/ak_subword.cuh"
#include "../../akshar_b200/csrc/ak_models.h"
#include "../../akshar_b200/csrc/unicode_tables.inc"

static AkTables host_tables() {
    AkTables T;
    T.page_index = ak_tbl_page_index; T.leaves = ak_tbl_leaves;
    T.decomp_keys = ak_tbl_decomp_keys; T.decomp_off = ak_tbl_decomp_off; T.decomp_data = ak_tbl_decomp_data;
    T.pair_keys = ak_tbl_pair_keys; T.pair_vals = ak_tbl_pair_vals;
    T.ll_keys = ak_tbl_latin_lower_keys; T.ll_vals = ak_tbl_latin_lower_vals;
    T.fl_keys = ak_tbl_full_lower_keys; T.fl_vals = ak_tbl_full_lower_vals;
    T.n_decomp = AK_N_DECOMP; T.n_pairs = AK_N_PAIRS; T.n_ll = AK_N_LATIN_LOWER; T.n_fl = AK_N_FULL_LOWER;
    return T;
}

extern "C" {

// spans: n_spans+1 absolute boundaries starting at off[0] and ending at off[n_rows]+1
int64_t hh_normalize(const uint8_t* text, const int64_t* off, int64_t n_rows, uint32_t flags, const int64_t* spans,
                     int64_t n_spans, int64_t limit, uint8_t* out, int64_t* out_off, uint32_t* status) {
    AkTables T = host_tables();
    std::vector<int64_t> cnt(n_spans + 1, 0);
    uint32_t st = 0;
    for (int64_t i = 0; i < n_spans; ++i)
        cnt[i + 1] = cnt[i] + ak_norm_span(T, text, off, n_rows, 0, n_rows, spans[i], spans[i + 1], flags, limit, nullptr,
                                           nullptr, 0, st);
    for (int64_t i = 0; i < n_spans; ++i) {
        int64_t c = ak_norm_span(T, text, off, n_rows, 0, n_rows, spans[i], spans[i + 1], flags, limit, out + cnt[i],
                                 out_off, cnt[i], st);
        if (c != cnt[i + 1] - cnt[i]) st |= 0x80000000u;
    }
    *status = st;
    return cnt[n_spans];
}

void hh_segment(const uint8_t* text, const int64_t* off, int64_t n_rows, uint32_t flags, const int64_t* spans,
                int64_t n_spans, int64_t limit, int32_t* cluster_ends, int64_t* cluster_splits, int32_t* run_ends,
                uint8_t* run_tags, int64_t* run_splits, int64_t cap, int64_t* totals, uint32_t* status) {
    AkTables T = host_tables();
    std::vector<int64_t> cc(n_spans + 1, 0), rc(n_spans + 1, 0);
    uint32_t st = 0;
    AkSegOut o;
    memset(&o, 0, sizeof(o));
    for (int64_t i = 0; i < n_spans; ++i) {
        int64_t a, b;
        ak_seg_span(T, text, off, n_rows, 0, n_rows, spans[i], spans[i + 1], flags, limit, false, o, a, b, st);
        cc[i + 1] = cc[i] + a;
        rc[i + 1] = rc[i] + b;
    }
    o.cluster_ends = cluster_ends; o.cluster_splits = cluster_splits; o.run_ends = run_ends; o.run_tags = run_tags;
    o.run_splits = run_splits; o.ccap = cap; o.rcap = cap;
    for (int64_t i = 0; i < n_spans; ++i) {
        int64_t a, b;
        o.cbase = cc[i]; o.rbase = rc[i];
        ak_seg_span(T, text, off, n_rows, 0, n_rows, spans[i], spans[i + 1], flags, limit, true, o, a, b, st);
    }
    totals[0] = cc[n_spans];
    totals[1] = rc[n_spans];
    *status = st;
}


// roman_phonetic_signature of every row; returns output bytes
int64_t hh_signature(const uint8_t* text, const int64_t* off, int64_t n_rows, uint8_t* out, int64_t* out_off) {
    AkTables T = host_tables();
    int64_t o = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        std::vector<uint32_t> a((size_t)(off[r + 1] - off[r]) + 1);
        int n = ak_signature_row(T, text, off[r], off[r + 1], a.data());
        out_off[r] = o;
        for (int i = 0; i < n; ++i) o += ak_encode(a[i], out + o);
    }
    out_off[n_rows] = o;
    return o;
}

static AkBpeHost g_bpe;
static AkUniHost g_uni;
static std::string g_err;
const char* hh_error() { return g_err.c_str(); }

int hh_load_bpe(const char* json, int64_t len) {
    g_bpe = AkBpeHost();
    g_err = ak_parse_bpe_json(json, (size_t)len, g_bpe);
    return g_err.empty() ? 0 : -1;
}
int hh_load_spm(const uint8_t* proto, int64_t len) {
    g_uni = AkUniHost();
    g_err = ak_parse_spm_model(proto, (size_t)len, g_uni);
    return g_err.empty() ? 0 : -1;
}
int hh_bpe_vocab_size() { return g_bpe.vocab_size; }
int hh_spm_vocab_size() { return (int)g_uni.piece.size(); }

// the kernel's structure, span by span: staged first walk, relative splits + fix-up, direct re-walk on overflow.
// returns total ids; *changed_out = NFC would change the text (caller re-runs on the NFC'd copy)
int64_t hh_bpe(const uint8_t* text, const int64_t* off, int64_t n_rows, const int64_t* spans, int64_t n_spans,
               int64_t limit, int stage_cap, int32_t* ids, int64_t id_cap, int64_t* splits, int* changed_out,
               uint32_t* status) {
    AkTables T = host_tables();
    AkBpeDev M;
    M.cp_direct = g_bpe.cp_direct.data(); M.cp_keys = g_bpe.cp_keys.data(); M.cp_ids = g_bpe.cp_ids.data();
    M.n_cp = (int)g_bpe.cp_keys.size(); M.mkeys = g_bpe.mkeys.data(); M.mvals = g_bpe.mvals.data(); M.mbits = g_bpe.mbits;
    M.bos = g_bpe.bos; M.eos = g_bpe.eos;
    std::vector<int32_t> poolbuf(1 << 20);
    unsigned long long used = 0;
    AkPool pool; pool.base = poolbuf.data(); pool.used = &used; pool.cap = poolbuf.size();
    uint32_t st = 0;
    bool changed = false;
    int64_t base = 0;
    std::vector<int32_t> stage((size_t)stage_cap + 1);
    for (int64_t i = 0; i < n_spans; ++i) {
        AkIdSink sink; sink.buf = stage.data(); sink.cap = stage_cap; sink.stride = 1; sink.cnt = 0; sink.direct = false;
        sink.gout = ids; sink.gbase = 0; sink.gcap = id_cap;
        int64_t rf, rl;
        ak_bpe_span(M, T, text, off, n_rows, 0, n_rows, spans[i], spans[i + 1], limit, sink, splits, 0, rf, rl, pool, changed, st);
        for (int64_t r = rf; r < rl; ++r) splits[r] += base;
        if (sink.cnt <= stage_cap) {
            for (int k = 0; k < sink.cnt; ++k) if (base + k < id_cap) ids[base + k] = stage[(size_t)k];
        } else {
            AkIdSink s2 = sink; s2.cnt = 0; s2.direct = true; s2.gbase = base;
            bool c2 = false; int64_t a, b;
            ak_bpe_span(M, T, text, off, n_rows, 0, n_rows, spans[i], spans[i + 1], limit, s2, nullptr, 0, a, b, pool, c2, st);
            if (s2.cnt != sink.cnt) st |= 0x80000000u;
        }
        base += sink.cnt;
    }
    *changed_out = changed ? 1 : 0;
    *status = st;
    return base;
}

int64_t hh_unigram(const uint8_t* text, const int64_t* off, int64_t n_rows, int32_t* ids, int64_t id_cap, int64_t* splits) {
    AkUniDev U;
    U.tkeys = g_uni.tkeys.data(); U.tvals = g_uni.tvals.data(); U.tbits = g_uni.tbits; U.score = g_uni.score.data();
    U.usable = g_uni.usable.data(); U.byte_id = g_uni.byte_id; U.unk_id = g_uni.unk_id; U.unk_score = g_uni.unk_score;
    U.flags = g_uni.flags;
    int64_t base = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        std::vector<uint32_t> back((size_t)(off[r + 1] - off[r]) + 3);
        int64_t n = ak_unigram_forward(U, text, off[r], off[r + 1], back.data());
        int64_t cnt = ak_unigram_backtrack(U, back.data(), n, nullptr, 0, 0);
        splits[r] = base;
        ak_unigram_backtrack(U, back.data(), n, ids, base + cnt, id_cap);
        base += cnt;
    }
    splits[n_rows] = base;
    return base;
}

}  // extern "C"
