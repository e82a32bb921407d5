// CPU build of the span walkers (akshar_b200/csrc/ak_text_core.cuh) -- TEST AID ONLY.
// It lets tests/test_span_walkers.py check, without a GPU, that cutting the buffer into arbitrary spans never
// changes the result (the property the CUDA kernels rely on).  Not linked into libakshar_b200.so.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../akshar_b200/csrc/ak_text_core.cuh"
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

}  // extern "C"
