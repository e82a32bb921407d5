// CPU build of the span walkers (akshar_b200/csrc/ak_text_core.cuh) -- TEST AID ONLY.
// It lets tests/test_span_walkers.py check, without a GPU, that cutting the buffer into arbitrary spans never
// changes the result (the property the CUDA kernels rely on).  Not linked into libakshar_b200.so.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <math.h>
#include "../../akshar_b200/csrc/ak_subword.cuh"
#include "../../akshar_b200/csrc/ak_fast.cuh"
#include "../../akshar_b200/csrc/ak_norm3.cuh"
#include "../../akshar_b200/csrc/ak_seg3.cuh"
#include "../../akshar_b200/csrc/ak_bpe3.cuh"
#include "../../akshar_b200/csrc/ak_tok.cuh"
#include "../../akshar_b200/csrc/ak_wordtok.cuh"
#include "../../akshar_b200/csrc/ak_lines.cuh"
#include "../../akshar_b200/csrc/ak_tok_host.h"
#include "../../akshar_b200/csrc/ak_decode_host.h"
#include "../../akshar_b200/csrc/ak_models.h"
#include "../../akshar_b200/csrc/unicode_tables.inc"

static AkTables host_tables() {
    AkTables T;
    T.page_index = ak_tbl_page_index; T.leaves = ak_tbl_leaves;
    T.decomp_keys = ak_tbl_decomp_keys; T.decomp_off = ak_tbl_decomp_off; T.decomp_data = ak_tbl_decomp_data;
    T.pair_keys = ak_tbl_pair_keys; T.pair_vals = ak_tbl_pair_vals;
    T.ll_keys = ak_tbl_latin_lower_keys; T.ll_vals = ak_tbl_latin_lower_vals;
    T.fl_keys = ak_tbl_full_lower_keys; T.fl_vals = ak_tbl_full_lower_vals;
    T.kmap_keys = ak_tbl_kmap_keys; T.kmap_off = ak_tbl_kmap_off; T.kmap_data = ak_tbl_kmap_data; T.n_kmap = AK_N_KMAP;
    T.hf_unknown = ak_tbl_hf_unknown; T.n_hf_unknown = AK_N_HF_UNKNOWN;
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


// The fast normalize kernel's structure on the CPU: 16-byte chunks, "warps" of `real` chunks + 2 halo chunks, the
// same phase functions (ak_fast.cuh), the walker for slow chunks.  n_slow counts the chunks that took the slow lane.
static void hh_make_chunk(const uint8_t* text, int64_t cs, int64_t tb, int64_t te, const std::vector<uint8_t>& rowstart,
                          int64_t base0, AkChunk& c) {
    c.w[0] = c.w[1] = c.w[2] = c.w[3] = c.w[4] = 0;
    c.own = 0;
    c.rows = 0;
    for (int i = 0; i < 20; ++i) {
        int64_t q = cs + i;
        if (q >= tb && q < te) {
            c.w[i >> 2] |= (uint32_t)text[q] << ((i & 3) * 8);
            if (i < 16) c.own |= 1u << i;
        }
        if (i < 16 && q >= base0 && q - base0 < (int64_t)rowstart.size() && rowstart[(size_t)(q - base0)]) c.rows |= 1u << i;
    }
}

// ---- bit-parallel normalize (ak_norm3.cuh): 32-byte lanes, `real` lanes + 2 halo lanes per "warp", the kernel's
// three exchange rounds, its emit masks fed to the same 16-byte writer, the walker for slow lanes.
// roles of one byte value (for the exhaustive predicate test): the lane is filled with that byte
uint32_t hh_n3_roles(uint32_t byte) {
    uint32_t x[8];
    for (int i = 0; i < 8; ++i) x[i] = byte * 0x01010101u;
    AkN3Lane L;
    memset(&L, 0, sizeof(L));
    L.own = 0xFFFFFFFFu;
    akn3_phase1(x, L);
    uint32_t r = 0;
    const uint32_t m[18] = {L.cont, L.K, L.NL, L.E0b, L.F0b, L.LXo, L.A4b, L.A5b, L.A6b, L.A7b, L.x9Fb, L.NKb, L.NIb, L.R2b, L.R3b, L.QNb, L.B6b, L.B7b};
    for (int i = 0; i < 18; ++i) {
        if (m[i] != 0u && m[i] != 0xFFFFFFFFu) return 0xFFFFFFFFu;       // must be uniform
        if (m[i]) r |= 1u << i;
    }
    if (L.PL[5] & 1u) r |= 1u << 18;     // lowered bit 5
    return r;
}
// planes of 32 arbitrary bytes (transpose test)
void hh_n3_planes(const uint8_t* b, uint32_t* P) {
    uint32_t x[8];
    memcpy(x, b, 32);
    akb_planes(x, P);
}

int64_t hh_fast_normalize3(const uint8_t* text, const int64_t* off, int64_t n_rows, int real, uint32_t nflags, uint8_t* out,
                           int64_t* out_off, uint32_t* status, int64_t* n_slow) {
    AkTables T = host_tables();
    const int64_t tb = off[0], te = off[n_rows], base0 = tb;
    std::vector<uint8_t> rowstart((size_t)(te - base0) + 128, 0);
    for (int64_t r = 0; r <= n_rows; ++r) rowstart[(size_t)(off[r] - base0)] = 1;
    const int64_t n_lanes = (te - base0 + 1 + 31) / 32;
    const uint32_t NFLAGS = nflags;
    const bool raw = !(nflags & AK_NORM_CLEAN);
    int64_t base = 0, row = 0, slow_cnt = 0;
    uint32_t st = 0;
    const int NL = real + 2;
    std::vector<AkN3Lane> lanes((size_t)NL);
    std::vector<uint32_t> lastk((size_t)NL), rest((size_t)NL);
    for (int64_t w0 = 0; w0 < n_lanes; w0 += real) {
        for (int l = 0; l < NL; ++l) {
            AkN3Lane& L = lanes[(size_t)l];
            memset(&L, 0, sizeof(L));
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            uint32_t x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 32; ++i) {
                const int64_t q = cs + i;
                if (q >= tb && q < te) { x[i >> 2] |= (uint32_t)text[q] << ((i & 3) * 8); L.own |= 1u << i; }
                if (q >= base0 && q - base0 < (int64_t)rowstart.size() && rowstart[(size_t)(q - base0)]) L.rows |= 1u << i;
            }
            akn3_phase1(x, L);
        }
        for (int l = 0; l < NL; ++l) {
            const uint32_t up1p = l > 0 ? lanes[(size_t)l - 1].up1 : 0u;
            akn3_phase2(lanes[(size_t)l], up1p, l + 1 < NL ? lanes[(size_t)l + 1].dn1 : 0u, raw);
        }
        for (int l = 0; l < NL; ++l) {
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            // lane 0 has no left neighbour: conservative carries (previous not inert, previous an accent, previous dropped)
            akn3_phase3(T, text, cs, te, lanes[(size_t)l], l > 0 ? lanes[(size_t)l - 1].up2 : AKN3_HALO_UP2, l + 1 < NL ? lanes[(size_t)l + 1].dn2 : 0u, raw);
            rest[(size_t)l] = akn3_gaps_local(text, cs, te, lanes[(size_t)l]);
            lastk[(size_t)l] = akn3_last_kept(text, cs, te, lanes[(size_t)l]);
        }
        for (int l = 0; l < NL; ++l) {
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            akn3_gaps_remote(text, cs, te, lanes[(size_t)l], rest[(size_t)l], l > 0 ? lastk[(size_t)l - 1] : 0u);
            akn3_phase3b(lanes[(size_t)l]);
        }
        for (int l = 1; l <= real; ++l) {
            AkN3Lane& L = lanes[(size_t)l];
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            uint32_t info[2] = {0, 0};
            const bool fast = akn3_phase4(L, lanes[(size_t)l - 1].up3, lanes[(size_t)l + 1].dn1, lanes[(size_t)l + 1].dn3, info[0], info[1]);
            for (int h = 0; h < 2; ++h) {
                const int64_t hs = cs + 16 * h;
                const int64_t ss = hs < tb ? tb : hs, se = hs + 16 > te + 1 ? te + 1 : hs + 16;
                if (ss >= se) continue;
                while (row <= n_rows && off[row] < ss) ++row;
                if (!fast && h == 0 && getenv("AKN3_REASONS")) {
                    static long cnt[10], tot;
                    for (int b = 0; b < 10; ++b) if (L.flags & (0x100u << b)) cnt[b]++;
                    if (++tot % 200 == 0)
                        fprintf(stderr, "slow lanes %ld: loop-kept %ld gap-eq %ld gap-unknown %ld gap-remote-eq %ld own-T %ld first-dep %ld next %ld prev-T %ld pend-gap %ld pend-nextT %ld\n",
                                tot, cnt[0], cnt[1], cnt[2], cnt[3], cnt[4], cnt[5], cnt[6], cnt[7], cnt[8], cnt[9]);
                }
                if (!fast) {
                    ++slow_cnt;
                    base += ak_norm_span(T, text, off, n_rows, 0, n_rows, ss, se, NFLAGS, 0, out + base, out_off, base, st);
                    while (row <= n_rows && off[row] < se) ++row;
                } else {
                    AkChunk c;
                    hh_make_chunk(text, hs, tb, te, rowstart, base0, c);
                    const uint32_t emit = info[h];
                    while (row <= n_rows && off[row] < se) {
                        int i = (int)(off[row] - hs);
                        int before = 0;
                        for (int b = 0; b < i; ++b) before += (emit >> b) & 1u;
                        out_off[row] = base + before;
                        ++row;
                    }
                    base += akf_write(c, emit, out + base);
                }
            }
        }
    }
    *status = st;
    *n_slow = slow_cnt;
    return base;
}


// ---- bit-parallel segmentation (ak_seg3.cuh): byte roles of one byte value, and the kernel's lane structure
uint32_t hh_s3_roles(uint32_t byte) {
    uint32_t x[8];
    for (int i = 0; i < 8; ++i) x[i] = byte * 0x01010101u;
    AkS3Lane L;
    memset(&L, 0, sizeof(L));
    L.own = 0xFFFFFFFFu;
    aks3_phase1(x, L);
    const uint32_t m[18] = {L.cont, L.E0b, L.A4b, L.A5b, L.X4b, L.S4b, L.C4b, L.X5b, L.S5b, L.C5b, L.LKb, L.M4b, L.M5b, L.CR, L.LF, L.CTL, L.ROM, L.WEAK};
    uint32_t r = 0;
    for (int i = 0; i < 18; ++i) {
        if (m[i] != 0u && m[i] != 0xFFFFFFFFu) return 0xFFFFFFFFu;
        if (m[i]) r |= 1u << i;
    }
    return r;
}

void hh_seg_fast3(const uint8_t* text, const int64_t* off, int64_t n_rows, uint32_t flags, int real, int32_t* cluster_ends,
                  int64_t* cluster_splits, int32_t* run_ends, uint8_t* run_tags, int64_t* run_splits, int64_t cap, int64_t* totals,
                  uint32_t* status, int64_t* n_slow) {
    AkTables T = host_tables();
    const bool want_c = (flags & AK_SEG_CLUSTERS) != 0, want_r = (flags & AK_SEG_RUNS) != 0, matras = (flags & AK_SEG_MATRAS) != 0;
    const int64_t tb = off[0], te = off[n_rows], base0 = tb;
    std::vector<uint8_t> rowstart((size_t)(te - base0) + 128, 0);
    for (int64_t r = 0; r <= n_rows; ++r) rowstart[(size_t)(off[r] - base0)] = 1;
    const int64_t n_lanes = (te - base0 + 1 + 31) / 32;
    const int NL = real + 2;
    std::vector<AkS3Lane> lanes((size_t)NL);
    int64_t cbase = 0, rbase = 0, nr = 0, slow_cnt = 0;
    uint32_t st = 0;
    AkSegOut o;
    memset(&o, 0, sizeof(o));
    o.cluster_ends = cluster_ends; o.cluster_splits = cluster_splits; o.run_ends = run_ends; o.run_tags = run_tags;
    o.run_splits = run_splits; o.ccap = cap; o.rcap = cap;
    for (int64_t w0 = 0; w0 < n_lanes; w0 += real) {
        for (int l = 0; l < NL; ++l) {
            AkS3Lane& L = lanes[(size_t)l];
            memset(&L, 0, sizeof(L));
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            uint32_t x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 32; ++i) {
                const int64_t q = cs + i;
                if (q >= tb && q < te) { x[i >> 2] |= (uint32_t)text[q] << ((i & 3) * 8); L.own |= 1u << i; }
                if (q >= base0 && q - base0 < (int64_t)rowstart.size() && rowstart[(size_t)(q - base0)]) L.rows |= 1u << i;
            }
            aks3_phase1(x, L);
        }
        for (int l = 0; l < NL; ++l) {
            aks3_phase2(lanes[(size_t)l], l + 1 < NL ? lanes[(size_t)l + 1].dn1 : 0u);
            if (lanes[(size_t)l].FOR) aks3_foreign(T, text, base0 + (w0 - 1 + l) * 32, te, lanes[(size_t)l]);
            aks3_summary(lanes[(size_t)l]);
        }
        for (int l = 1; l <= real; ++l) {
            AkS3Lane& L = lanes[(size_t)l];
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            const int64_t ss = cs < tb ? tb : cs, se = cs + 32 > te + 1 ? te + 1 : cs + 32;
            if (ss >= se) continue;
            while (nr <= n_rows && off[nr] < ss) ++nr;
            const uint32_t tb_bit = (tb >= cs && tb < cs + 32) ? 1u << (int)(tb - cs) : 0u;
            const bool fast = aks3_phase3(L, lanes[(size_t)l - 1].up2, tb_bit, matras, want_c, want_r);
            if (!fast) {
                ++slow_cnt;
                int64_t a = 0, b = 0;
                o.cbase = cbase; o.rbase = rbase;
                ak_seg_span(T, text, off, n_rows, 0, n_rows, ss, se, flags, 0, true, o, a, b, st);
                cbase += a; rbase += b;
                while (nr <= n_rows && off[nr] < se) ++nr;
            } else {
                const uint32_t rows_ev = L.rows & ~tb_bit;
                const uint32_t mc = want_c ? (L.brk | rows_ev) : 0u, mr = want_r ? (L.rchg | rows_ev) : 0u;
                const int64_t rs_in = nr > 0 ? off[nr - 1] : off[0];
                if (want_c) aks3_emit(L, mc, cs, rs_in, cluster_ends + cbase, nullptr);
                if (want_r) aks3_emit(L, mr, cs, rs_in, run_ends + rbase, run_tags + rbase);
                nr = aks3_splits(L, mc, mr, cs, off, n_rows, nr, cbase, rbase, want_c ? cluster_splits : nullptr, want_r ? run_splits : nullptr);
                cbase += akb_popc(mc);
                rbase += akb_popc(mr);
            }
        }
    }
    totals[0] = cbase;
    totals[1] = rbase;
    *status = st;
    *n_slow = slow_cnt;
}

// word tokenizers (ak_wordtok.cuh): the kernel's lane structure -- `real` lanes + 2 halo lanes per "warp", count pass, prefix,
// emit pass through the same akwt_emit_lane
int64_t hh_wordtok(const uint8_t* text, const int64_t* off, int64_t n_rows, int mode, int real, int32_t* begin, int32_t* end,
                   int64_t cap, int64_t* splits, uint8_t* row_flags, uint32_t* status) {
    const int64_t tb = off[0], te = off[n_rows], base0 = tb;
    std::vector<uint8_t> rowstart((size_t)(te - base0) + 128, 0);
    for (int64_t r = 0; r <= n_rows; ++r) rowstart[(size_t)(off[r] - base0)] = 1;
    const int64_t n_lanes = (te - base0 + 1 + 31) / 32;
    const int NL = real + 2;
    std::vector<AkWtLane> lanes((size_t)NL);
    std::vector<uint32_t> open((size_t)NL);
    uint32_t st = 0;
    if (row_flags) memset(row_flags, 0, (size_t)n_rows);
    int64_t total = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int64_t t_at = 0, nr = 0;
        for (int64_t w0 = 0; w0 < n_lanes; w0 += real) {
            for (int l = 0; l < NL; ++l) {
                AkWtLane& L = lanes[(size_t)l];
                memset(&L, 0, sizeof(L));
                const int64_t cs = base0 + (w0 - 1 + l) * 32;
                uint32_t x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int i = 0; i < 32; ++i) {
                    const int64_t q = cs + i;
                    if (q >= tb && q < te) { x[i >> 2] |= (uint32_t)text[q] << ((i & 3) * 8); L.own |= 1u << i; }
                    if (q == te) L.endbit = 1u << i;
                    if (q >= base0 && q - base0 < (int64_t)rowstart.size() && rowstart[(size_t)(q - base0)]) L.rows |= 1u << i;
                }
                akwt_phase1(x, L);
            }
            for (int l = 0; l < NL; ++l) {
                AkWtLane& L = lanes[(size_t)l];
                akwt_phase2(L, l + 1 < NL ? lanes[(size_t)l + 1].dn : 0u, mode);
                if (L.hl & ~L.DEV) akwt_wide(text, base0 + (w0 - 1 + l) * 32, te, L);
                akwt_summary(L);
            }
            for (int l = 0; l < NL; ++l) open[(size_t)l] = akwt_phase3(lanes[(size_t)l], l > 0 ? lanes[(size_t)l - 1].up : 0u);
            for (int l = 1; l <= real; ++l) {
                const AkWtLane& L = lanes[(size_t)l];
                const int64_t cs = base0 + (w0 - 1 + l) * 32;
                if (pass == 1) {
                    while (nr <= n_rows && off[nr] < cs) ++nr;
                    int nrows = 0;
                    while (nr + nrows <= n_rows && off[nr + nrows] < cs + 32) ++nrows;
                    akwt_emit_lane(L, cs, off, n_rows, nr, nrows, nr - 1, t_at, t_at - (int64_t)open[(size_t)l], begin, end, cap, splits,
                                   row_flags, st);
                }
                t_at += akb_popc(L.T);
            }
        }
        total = t_at;
    }
    *status = st;
    return total;
}

// file bytes -> rows (ak_lines.cuh) with the kernels' structure: per-thread transducers over `span`-byte stretches, composed
// in order, then the emit pass with the entry states; returns the number of rows
int64_t hh_lines(const uint8_t* text, int64_t n, int span, int64_t* begin, int64_t* end, int64_t cap, uint32_t* status) {
    uint32_t st = 0;
    const int64_t n_spans = n / span + 1;
    std::vector<AkLineFn> fn((size_t)n_spans);
    for (int64_t k = 0; k < n_spans; ++k) fn[(size_t)k] = akl_span(text, k * span, (k + 1) * span, n, false, 0u, -1, 0, nullptr, nullptr, 0, st);
    AkLineFn acc = akl_identity();
    int64_t rows = 0;
    for (int64_t k = 0; k < n_spans; ++k) {
        const uint32_t state = acc.s & 1u;                      // the file starts in state 0
        akl_span(text, k * span, (k + 1) * span, n, true, state, acc.lastk, acc.cnt0, begin, end, cap, st);
        acc = akl_compose(acc, fn[(size_t)k]);
    }
    rows = acc.cnt0;
    *status = st;
    return rows;
}

// the same through the kernels' 32-byte mask lanes (akln_*): `lead` bytes of padding in front make the lanes start before
// the file, as they do when the file's address is not 16-byte aligned
int64_t hh_lines32(const uint8_t* text, int64_t n, int lead, int64_t* begin, int64_t* end, int64_t cap, uint32_t* status) {
    uint32_t st = 0;
    const int64_t base0 = -(int64_t)lead;
    const int64_t n_lanes = (n - base0) / 32 + 1;
    std::vector<AkLnLane> L((size_t)n_lanes);
    std::vector<AkLineFn> fn((size_t)n_lanes);
    for (int64_t k = 0; k < n_lanes; ++k) {
        const int64_t cs = base0 + 32 * k;
        uint32_t x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        AkLnLane& l = L[(size_t)k];
        memset(&l, 0, sizeof(l));
        for (int i = 0; i < 32; ++i) {
            const int64_t q = cs + i;
            if (q >= 0 && q < n) { x[i >> 2] |= (uint32_t)text[q] << ((i & 3) * 8); l.own |= 1u << i; }
            if (q == n) l.endbit = 1u << i;
        }
        akln_phase1(x, l);
        if (l.WIDE) akln_wide(text, cs, n, l);
        fn[(size_t)k] = akln_summary(l, text, cs, n);
    }
    AkLineFn acc = akl_identity();
    for (int64_t k = 0; k < n_lanes; ++k) {
        akln_emit(L[(size_t)k], text, base0 + 32 * k, n, acc.s & 1u, acc.lastk, acc.cnt0, begin, end, cap, st);
        acc = akl_compose(acc, fn[(size_t)k]);
    }
    *status = st;
    return acc.cnt0;
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
// ids -> text (ak_decode.cuh) with the kernels' steps: marks row by row, lengths, prefix, bytes, row offsets
int64_t hh_decode(int kind, int form, const int32_t* ids, int64_t n_ids, const int64_t* splits, int64_t n_rows, uint8_t* out, int64_t cap,
                  int64_t* out_off, uint32_t* status) {
    const AkDecHost H = kind == 0 ? ak_build_bpe_decode(g_bpe, form) : ak_build_spm_decode(g_uni, form);
    const AkDecTable D = H.view();
    std::vector<uint8_t> mark((size_t)n_ids + 1, 0);
    uint32_t st = 0;
    for (int64_t r = 0; r < n_rows; ++r) akd_mark_row(D, form, ids, splits[r], splits[r + 1], mark.data(), st);
    std::vector<int64_t> at((size_t)n_ids + 1, 0);
    for (int64_t i = 0; i < n_ids; ++i) at[(size_t)i + 1] = at[(size_t)i] + akd_emit(D, ids, mark.data(), n_ids, i, (uint8_t*)nullptr, st);
    for (int64_t i = 0; i < n_ids; ++i) {
        const int64_t len = at[(size_t)i + 1] - at[(size_t)i];
        if (!len) continue;
        if (at[(size_t)i] + len <= cap) akd_emit(D, ids, mark.data(), n_ids, i, out + at[(size_t)i], st);
        else st |= AK_ST_OVERFLOW;
    }
    for (int64_t r = 0; r <= n_rows; ++r) out_off[r] = at[(size_t)splits[r]];
    *status = st;
    return at[(size_t)n_ids];
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


// ---- the event-stream encoders (ak_tok.cuh) on the CPU: the words kernel's lanes ("warps" of `real` lanes + 2 halo lanes,
// each with its block of `cap` event slots), then the row-fix, resolve, check and emit cores slot by slot in stream order.
// kind 0 BPE, 1 Unigram.  cache_bits: size of the word cache (small tables exercise probing and uncached words);
// prewarm = 0 starts it empty.  stats: [events, flagged rows, words sent to the exact Viterbi, cache misses, slots needed]
int64_t hh_tok(int kind, const uint8_t* text, const int64_t* off, int64_t n_rows, int real, int cap, int cache_bits, int prewarm,
               int u16, int splits_i32, void* ids, int64_t id_cap, void* splits, uint32_t* status, int64_t* stats) {
    AkTables T = host_tables();
    const int64_t tb = off[0], te = off[n_rows], base0 = tb;
    std::vector<uint8_t> rowstart((size_t)(te - base0) + 128, 0);
    for (int64_t r = 0; r <= n_rows; ++r) rowstart[(size_t)(off[r] - base0)] = 1;
    const int64_t n_lanes = (te - base0 + 1 + 31) / 32;
    const int64_t n_wt = (n_lanes + real - 1) / real;
    const int NL = real + 2;
    std::vector<AkEvent> slots((size_t)n_wt * cap);
    std::vector<uint32_t> count((size_t)n_wt, 0);
    std::vector<uint8_t> row_flag((size_t)n_rows + 2, 0);
    std::vector<uint32_t> row_ev((size_t)n_rows + 2, 0);
    int64_t result[4] = {0, 0, 0, 0};
    uint32_t st = 0;
    std::vector<AkB3Lane> lb((size_t)NL);
    std::vector<AkU3Lane> lu((size_t)NL);
    int64_t need = 0, n_events = 0;
    for (int64_t wt = 0; wt < n_wt; ++wt) {
        const int64_t w0 = wt * real;
        for (int l = 0; l < NL; ++l) {
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            uint32_t x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            uint32_t own = 0, rows = 0;
            for (int i = 0; i < 32; ++i) {
                const int64_t q = cs + i;
                if (q >= tb && q < te) { x[i >> 2] |= (uint32_t)text[q] << ((i & 3) * 8); own |= 1u << i; }
                if (q >= base0 && q - base0 < (int64_t)rowstart.size() && rowstart[(size_t)(q - base0)]) rows |= 1u << i;
            }
            if (kind == 0) {
                AkB3Lane& L = lb[(size_t)l];
                memset(&L, 0, sizeof(L));
                L.own = own; L.rows = rows;
                akb3_phase1(x, L);
            } else {
                AkU3Lane& L = lu[(size_t)l];
                memset(&L, 0, sizeof(L));
                L.own = own; L.rows = rows;
                aku3_phase1(x, L);
            }
        }
        if (kind == 0) {
            for (int l = 0; l < NL; ++l) {
                AkB3Lane& L = lb[(size_t)l];
                const int64_t cs = base0 + (w0 - 1 + l) * 32;
                akb3_phase2(L, l + 1 < NL ? lb[(size_t)l + 1].dn1 : 0u);
                if (L.FOR) akb3_foreign(T, text, cs, te, L);
                akb3_summary(L);
            }
            for (int l = 0; l < NL; ++l) akb3_phase3(lb[(size_t)l], l > 0 ? lb[(size_t)l - 1].up2 : 0u);
            lb[(size_t)NL - 1].bnd &= 0x3FFFFFFFu;
        } else {
            for (int l = 0; l < NL; ++l) aku3_phase2(lu[(size_t)l], l > 0 ? lu[(size_t)l - 1].up : 0u);
        }
        int pre = 0;
        for (int l = 1; l <= real; ++l) {
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            const int64_t ss = cs < tb ? tb : cs, se = cs + 32 > te + 1 ? te + 1 : cs + 32;
            if (ss >= se) continue;
            uint32_t rowsm, wstart, cw, bnd, nb1, nb2;
            if (kind == 0) {
                AkB3Lane& L = lb[(size_t)l];
                {
                    uint32_t fix = L.UNS & L.own;
                    const AkBpeDev hm = ak_bpe_host_view(g_bpe);
                    for (uint32_t m = L.LT & L.own; m;) {
                        const int i = akb_ctz(m);
                        m &= m - 1u;
                        if (ak_bpe_special_at(hm, text, cs + i, te) >= 0) fix |= 1u << i;
                    }
                    if (fix) ake_flag_rows(off, n_rows, cs, fix, row_flag.data());
                }
                if (L.trb && akb3_changes(T, text, off, n_rows, 0, L.trb, cs, st)) ake_flag_rows(off, n_rows, cs, L.trb, row_flag.data());
                rowsm = L.rows; wstart = L.wstart; cw = L.CW; bnd = L.bnd;
                nb1 = lb[(size_t)l + 1].bnd;
                nb2 = l + 2 < NL ? lb[(size_t)l + 2].bnd : 0u;
            } else {
                AkU3Lane& L = lu[(size_t)l];
                if (L.exotic) ake_flag_rows(off, n_rows, cs, L.exotic, row_flag.data());
                rowsm = L.rows; wstart = L.wstart; cw = 0xFFFFFFFFu; bnd = L.bnd;
                nb1 = lu[(size_t)l + 1].bnd;
                nb2 = l + 2 < NL ? lu[(size_t)l + 2].bnd : 0u;
            }
            // what akn3_lane_rows2 hands the lane: the first row that starts in its 32 bytes and how many do
            const int64_t first_row = ak_row_lower_bound(off, 0, n_rows, cs);
            int nrows = 0;
            for (int64_t g = first_row; g <= n_rows && off[g] < cs + 32; ++g) ++nrows;
            const int tail = l >= real ? 62 : l == real - 1 ? 94 : 96;
            const int64_t at = wt * cap + pre;
            pre += ake_lane_events(rowsm, wstart, cw, bnd, nb1, nb2, tail, cs, tb, off, n_rows, first_row, nrows, at, row_ev.data(),
                                   slots.data() + at, (int64_t)cap - pre,
                                   [&](int64_t p, int64_t from) -> int64_t {
                                       if (kind == 0) return akb3_scan_end(T, text, p, from, (cw >> (int)(p - cs)) & 1u, off, n_rows, 0, n_rows);
                                       const int64_t er = ak_row_lower_bound(off, 0, n_rows, p + 1);
                                       const int64_t re = off[er];
                                       int64_t q = from > re ? re : from;
                                       while (q < re && text[q] != 0x20u) ++q;
                                       return q;
                                   });
        }
        count[(size_t)wt] = (uint32_t)(pre < cap ? pre : cap);
        if (pre > cap) st |= AK_ST_OVERFLOW;
        if (pre > need) need = pre;
        n_events += pre;
    }
    if (kind == 1)
        for (int64_t g = 0; g < n_rows; ++g)
            if (off[g + 1] - off[g] > 8192) row_flag[(size_t)g] = 1;       // ak_long_rows_kernel
    stats[0] = n_events;
    stats[4] = need;
    if (st & AK_ST_OVERFLOW) { *status = st; stats[1] = stats[2] = stats[3] = 0; return 0; }
    // model + cache
    AkTokModel M;
    memset(&M, 0, sizeof(M));
    M.kind = kind;
    M.T = T;
    std::vector<unsigned long long> img;
    if (kind == 0) {
        M.bpe = ak_bpe_host_view(g_bpe);
        img = prewarm ? ak_build_bpe_image(g_bpe, T, (uint32_t)cache_bits) : std::vector<unsigned long long>((size_t)AKC_ENTRY << cache_bits, 0ull);
    } else {
        M.uni = ak_uni_host_view(g_uni);
        if (!ak_uni_wordwise(g_uni)) { *status = 0x40000000u; return -1; }
        img = prewarm ? ak_build_uni_image(g_uni, (uint32_t)cache_bits) : std::vector<unsigned long long>((size_t)AKC_ENTRY << cache_bits, 0ull);
    }
    M.cache.e = img.data();
    M.cache.bits = (uint32_t)cache_bits;
    M.cache.inserted = nullptr;
    std::vector<int32_t> longpool(1 << 20), pool((size_t)(te - tb) * 3 + (1 << 20));
    unsigned long long long_used = 0, pool_used = 0;
    M.pool.base = longpool.data(); M.pool.used = &long_used; M.pool.cap = longpool.size();
    std::vector<unsigned long long> row_fix((size_t)n_rows + 1, 0ull);
    // row fix
    AkRowFixCtx RX;
    RX.M = M; RX.text = text; RX.off = off; RX.n_rows = n_rows; RX.result = result; RX.ev = slots.data();
    RX.n_events = slots.size(); RX.row_ev = row_ev.data(); RX.row_fix = row_fix.data();
    RX.pool = pool.data(); RX.pool_used = &pool_used; RX.pool_cap = pool.size();
    int64_t n_flagged = 0;
    for (int64_t g = 0; g < n_rows; ++g) if (row_flag[(size_t)g]) { akr_fix_row(RX, g); ++n_flagged; }
    AkLookupCtx X;
    X.M = M; X.text = text; X.off = off; X.n_rows = n_rows; X.tb = tb; X.te = te; X.result = result;
    X.ids = ids; X.id_cap = id_cap; X.ids_u16 = u16; X.splits = splits; X.splits_i32 = splits_i32;
    X.row_flag = row_flag.data(); X.row_fix = row_fix.data(); X.pool = pool.data(); X.pool_used = &pool_used; X.pool_cap = pool.size();
    X.any_fix = n_flagged ? 1 : 0;
    // resolve
    std::vector<unsigned long long> resolved(slots.size(), 0ull);
    std::vector<uint32_t> aux(slots.size(), 0u);
    int64_t n_miss = 0, n_exact = 0;
    for (int64_t wt = 0; wt < n_wt; ++wt)
        for (uint32_t o = 0; o < count[(size_t)wt]; ++o) {
            const size_t s = (size_t)wt * cap + o;
            AkEvent& ev = slots[s];
            const uint32_t k = ev.meta & 7u, len = ev.meta >> 3;
            unsigned long long kw[4] = {0, 0, 0, 0};
            if (k <= AKE_WORD && len <= AKC_MAXLEN) akc_key0123(text, tb + ev.pos, len, te, kw);
            const unsigned long long before = long_used + pool_used;
            resolved[s] = kind == 0 ? akl_resolve<0>(X, ev, kw, aux[s], st) : akl_resolve<1>(X, ev, kw, aux[s], st);
            if (k <= AKE_WORD && (resolved[s] >> 62) != AKR_CACHE && long_used + pool_used != before) ++n_miss;
        }
    // check (Unigram)
    if (kind == 1) {
        float d = 0.f;
        for (int64_t wt = 0; wt < n_wt; ++wt)
            for (uint32_t o = 0; o < count[(size_t)wt]; ++o) {
                const size_t s = (size_t)wt * cap + o;
                if ((resolved[s] >> 62) == AKR_EVENT) { d = 0.f; continue; }
                if (aux[s] == 0u) continue;
                d += aku_aux_wmag(aux[s]);
                if (!aku_robust(aku_aux_ratio(aux[s]), d)) {
                    const AkEvent ev = slots[s];
                    long long pa;
                    const int n = akl_uni_exact(X, tb + ev.pos, ev.meta >> 3, &pa);
                    resolved[s] = pa >= 0 ? akr_pool(n, (unsigned long long)pa) : 0ull;
                    ++n_exact;
                }
            }
    }
    // emit
    int64_t at = 0;
    for (int64_t wt = 0; wt < n_wt; ++wt)
        for (uint32_t o = 0; o < count[(size_t)wt]; ++o) {
            const size_t s = (size_t)wt * cap + o;
            const int n = akr_n(resolved[s]);
            if (n && at + n > id_cap) st |= AK_ST_OVERFLOW;
            if (resolved[s]) akl_emit(X, resolved[s], &slots[s], at);
            at += n;
        }
    stats[1] = n_flagged;
    stats[2] = n_exact;
    stats[3] = n_miss;
    *status = st | (uint32_t)result[2];
    return at;
}

int64_t hh_unigram(const uint8_t* text, const int64_t* off, int64_t n_rows, int32_t* ids, int64_t id_cap, int64_t* splits) {
    AkUniDev U;
    U.tkeys = g_uni.tkeys.data(); U.tvals = g_uni.tvals.data(); U.tkv = nullptr; U.tbits = g_uni.tbits; U.score = g_uni.score.data();
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

