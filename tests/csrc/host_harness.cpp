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
#include "../../akshar_b200/csrc/ak_bpe_fast.cuh"
#include "../../akshar_b200/csrc/ak_seg_fast.cuh"
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

int64_t hh_fast_normalize(const uint8_t* text, const int64_t* off, int64_t n_rows, int real, uint8_t* out, int64_t* out_off,
                          uint32_t* status, int64_t* n_slow) {
    AkTables T = host_tables();
    std::vector<uint32_t> lut(384);
    for (int i = 0; i < 384; ++i) lut[(size_t)i] = i < 128 ? ak_props(T, (uint32_t)i) : ak_props(T, 0x900u + (uint32_t)(i - 128));
    const int64_t tb = off[0], te = off[n_rows], base0 = tb;
    std::vector<uint8_t> rowstart((size_t)(te - base0) + 64, 0);
    for (int64_t r = 0; r <= n_rows; ++r) rowstart[(size_t)(off[r] - base0)] = 1;
    const int64_t n_chunks = (te - base0 + 1 + 15) / 16;
    const uint32_t NFLAGS = AK_NORM_ROMAN | AK_NORM_CLEAN;
    int64_t base = 0, row = 0, slow_cnt = 0;
    uint32_t st = 0;
    std::vector<AkChunk> lanes((size_t)real + 2);
    for (int64_t w0 = 0; w0 < n_chunks; w0 += real) {
        for (int l = 0; l < real + 2; ++l) {
            AkChunk& c = lanes[(size_t)l];
            int64_t cs = base0 + (w0 - 1 + l) * 16;
            hh_make_chunk(text, cs, tb, te, rowstart, base0, c);
            if (c.own == 0 && c.rows == 0) {
                c.kept = c.lead = 0;
                c.flags = AKF_BOUNDARY | AKF_ROWSTART;
                c.first_w = c.last_w = c.F = c.L1 = c.L2 = AKF_NONE;
            } else {
                akf_phase_a(T, lut.data(), c);
            }
        }
        for (int l = 0; l < real + 2; ++l) {
            AkChunk& c = lanes[(size_t)l];
            if (l == 0) {
                // left halo: the kernel decodes the code point that ends right before the chunk
                uint32_t pl = AKF_NONE;
                int64_t cs = base0 + (w0 - 1) * 16;
                if (c.first_w != AKF_NONE && cs > tb) {
                    int64_t q = cs - 1;
                    int k = 0;
                    while (q > tb && k < 3 && (text[q] & 0xC0u) == 0x80u) { --q; ++k; }
                    int len;
                    pl = akf_props(T, lut.data(), ak_decode(text, q, te, len));
                }
                akf_resolve_first(c, pl);
            } else {
                akf_resolve_first(c, lanes[(size_t)l - 1].last_w);
            }
        }
        for (int l = 1; l <= real; ++l) {
            AkChunk& c = lanes[(size_t)l];
            int64_t cs = base0 + (w0 - 1 + l) * 16;
            int64_t ss = cs < tb ? tb : cs, se = cs + 16 > te + 1 ? te + 1 : cs + 16;
            if (ss >= se) continue;
            AkNeighbor pv, nx;
            pv.flags = lanes[(size_t)l - 1].flags; pv.F = AKF_NONE; pv.L1 = lanes[(size_t)l - 1].L1; pv.L2 = lanes[(size_t)l - 1].L2;
            nx.flags = lanes[(size_t)l + 1].flags; nx.F = lanes[(size_t)l + 1].F; nx.L1 = nx.L2 = AKF_NONE;
            uint32_t emit = 0;
            bool slow = akf_is_slow(c, pv, nx) || !akf_collapse(c, pv, nx, emit);
            if (slow && getenv("AKF_REASONS")) {
                int why = (c.flags & AKF_TROUBLE) ? 0 : ((c.flags & AKF_FIRST_DEP) && ((pv.flags & AKF_TROUBLE) || !(pv.flags & AKF_BOUNDARY))) ? 1
                        : ((nx.flags & AKF_LEAD_TROUBLE) || !(nx.flags & AKF_BOUNDARY)) ? 2
                        : ((c.flags & AKF_KEPT_BEFORE_ROW) && ((pv.flags & AKF_TROUBLE) || (!(pv.flags & AKF_ROWSTART) && pv.L2 == AKF_NONE))) ? 3 : 4;
                static long cnts[5];
                cnts[why]++;
                if ((cnts[0] + cnts[1] + cnts[2] + cnts[3] + cnts[4]) % 500 == 0)
                    fprintf(stderr, "reasons own-trouble %ld first-dep %ld next %ld prev-kept %ld lookahead %ld\n", cnts[0], cnts[1], cnts[2], cnts[3], cnts[4]);
                if (why == 0 && cnts[0] < 12) {
                    fprintf(stderr, "  trouble chunk: ");
                    for (int i = 0; i < 19; ++i) fprintf(stderr, "%02x ", akf_byte(c, i));
                    fprintf(stderr, "\n");
                }
            }
            while (row <= n_rows && off[row] < ss) ++row;
            if (slow) {
                ++slow_cnt;
                base += ak_norm_span(T, text, off, n_rows, 0, n_rows, ss, se, NFLAGS, 0, out + base, out_off, base, st);
                while (row <= n_rows && off[row] < se) ++row;
            } else {
                while (row <= n_rows && off[row] < se) {
                    int i = (int)(off[row] - cs);
                    int before = 0;
                    for (int b = 0; b < i; ++b) before += (emit >> b) & 1u;
                    out_off[row] = base + before;
                    ++row;
                }
                base += akf_write(c, emit, out + base);
            }
        }
    }
    *status = st;
    *n_slow = slow_cnt;
    return base;
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

// The fast BPE kernel's structure on the CPU: chunks, halo lanes, word cache (starting empty, `cache_bits` slots so
// that small tables exercise probing / eviction-free misses), lane emit with relative splits.
int64_t hh_bpe_fast(const uint8_t* text, const int64_t* off, int64_t n_rows, int real, int cache_bits, int stage_cap,
                    int32_t* ids, int64_t id_cap, int64_t* splits, int* changed_out, uint32_t* status) {
    AkTables T = host_tables();
    std::vector<uint32_t> lut(384);
    for (int i = 0; i < 384; ++i) lut[(size_t)i] = i < 128 ? ak_props(T, (uint32_t)i) : ak_props(T, 0x900u + (uint32_t)(i - 128));
    AkBpeDev M;
    M.cp_direct = g_bpe.cp_direct.data(); M.cp_keys = g_bpe.cp_keys.data(); M.cp_ids = g_bpe.cp_ids.data();
    M.n_cp = (int)g_bpe.cp_keys.size(); M.mkeys = g_bpe.mkeys.data(); M.mvals = g_bpe.mvals.data(); M.mbits = g_bpe.mbits;
    M.bos = g_bpe.bos; M.eos = g_bpe.eos;
    std::vector<unsigned long long> img((size_t)AKW_ENTRY << cache_bits, 0ull);
    AkWordCache C; C.e = img.data(); C.bits = (uint32_t)cache_bits;
    std::vector<int32_t> poolbuf(1 << 20);
    unsigned long long used = 0;
    AkPool pool; pool.base = poolbuf.data(); pool.used = &used; pool.cap = poolbuf.size();
    const int64_t tb = off[0], te = off[n_rows], base0 = tb;
    std::vector<uint8_t> rowstart((size_t)(te - base0) + 64, 0);
    for (int64_t r = 0; r <= n_rows; ++r) rowstart[(size_t)(off[r] - base0)] = 1;
    const int64_t n_chunks = (te - base0 + 1 + 15) / 16;
    AkBLaneCtx X; X.M = &M; X.T = &T; X.C = &C; X.text = text; X.off = off; X.n_rows = n_rows; X.r_lo = 0; X.r_hi = n_rows; X.pool = &pool;
    uint32_t st = 0;
    bool changed = false;
    int64_t base = 0;
    std::vector<AkBChunk> lanes((size_t)real + 2);
    std::vector<int32_t> stage((size_t)stage_cap + 1);
    for (int64_t w0 = 0; w0 < n_chunks; w0 += real) {
        for (int l = 0; l < real + 2; ++l) {
            AkBChunk& c = lanes[(size_t)l];
            int64_t cs = base0 + (w0 - 1 + l) * 16;
            AkChunk tmp;
            hh_make_chunk(text, cs, tb, te, rowstart, base0, tmp);
            for (int k = 0; k < 5; ++k) c.w[k] = tmp.w[k];
            c.rows = tmp.rows; c.own = tmp.own;
            akb_phase_a(T, lut.data(), c);
        }
        for (int l = 0; l < real + 2; ++l) {
            AkBChunk& c = lanes[(size_t)l];
            uint32_t pw = AKF_NONE, pk = 2;
            if (l == 0) {
                int64_t cs = base0 + (w0 - 1) * 16;
                if (c.first_pos < 32u && cs > tb) {
                    int64_t q = cs - 1;
                    int k = 0;
                    while (q > tb && k < 3 && (text[q] & 0xC0u) == 0x80u) { --q; ++k; }
                    int len;
                    pw = akf_props(T, lut.data(), ak_decode(text, q, te, len));
                    pk = AK_HFCLASS(pw);
                }
            } else {
                pw = lanes[(size_t)l - 1].last_w; pk = lanes[(size_t)l - 1].last_cls;
            }
            akb_resolve_first(c, pw, pk);
        }
        for (int l = 1; l <= real; ++l) {
            AkBChunk& c = lanes[(size_t)l];
            int64_t cs = base0 + (w0 - 1 + l) * 16;
            int64_t ss = cs < tb ? tb : cs, se = cs + 16 > te + 1 ? te + 1 : cs + 16;
            if (ss >= se) continue;
            if (c.flags & AKB_ALPHABET) st |= AK_ST_ALPHABET;
            if ((c.flags & AKF_TROUBLE) && akb_chunk_changes(X, c, cs, 0, st)) changed = true;
            AkIdSink sink; sink.buf = stage.data(); sink.cap = stage_cap; sink.stride = 1; sink.cnt = 0; sink.direct = false;
            sink.gout = ids; sink.gbase = 0; sink.gcap = id_cap;
            int64_t rf, rl;
            const uint32_t nb = (lanes[(size_t)l + 1].bnd & 0xFFFFu) | (l + 2 <= real + 1 && l < real ? (lanes[(size_t)l + 2].bnd & 0xFFFFu) << 16 : 0u);
            akb_lane_emit(X, c, nb, cs, sink, splits, rf, rl, st);
            for (int64_t r = rf; r < rl; ++r) splits[r] += base;
            if (sink.cnt <= stage_cap) {
                for (int k = 0; k < sink.cnt; ++k) if (base + k < id_cap) ids[base + k] = stage[(size_t)k];
            } else {
                AkIdSink s2 = sink; s2.cnt = 0; s2.direct = true; s2.gbase = base;
                int64_t a, b;
                akb_lane_emit(X, c, nb, cs, s2, nullptr, a, b, st);
                if (s2.cnt != sink.cnt) st |= 0x80000000u;
            }
            base += sink.cnt;
        }
    }
    *changed_out = changed ? 1 : 0;
    *status = st;
    return base;
}


// ---- bit-parallel BPE front end (ak_bpe3.cuh): lanes of 32 bytes, boundaries / word starts / trouble bits from the
// planes, every word through the exact merge loop (the word cache is the kernel's business)
int64_t hh_bpe_fast3(const uint8_t* text, const int64_t* off, int64_t n_rows, int real, int32_t* ids, int64_t id_cap,
                     int64_t* splits, int* changed_out, uint32_t* status) {
    AkTables T = host_tables();
    AkBpeDev M;
    M.cp_direct = g_bpe.cp_direct.data(); M.cp_keys = g_bpe.cp_keys.data(); M.cp_ids = g_bpe.cp_ids.data();
    M.n_cp = (int)g_bpe.cp_keys.size(); M.mkeys = g_bpe.mkeys.data(); M.mvals = g_bpe.mvals.data(); M.mbits = g_bpe.mbits;
    M.bos = g_bpe.bos; M.eos = g_bpe.eos;
    std::vector<int32_t> poolbuf(1 << 20);
    unsigned long long used = 0;
    AkPool pool; pool.base = poolbuf.data(); pool.used = &used; pool.cap = poolbuf.size();
    const int64_t tb = off[0], te = off[n_rows], base0 = tb;
    std::vector<uint8_t> rowstart((size_t)(te - base0) + 128, 0);
    for (int64_t r = 0; r <= n_rows; ++r) rowstart[(size_t)(off[r] - base0)] = 1;
    const int64_t n_lanes = (te - base0 + 1 + 31) / 32;
    const int NL = real + 2;
    std::vector<AkB3Lane> lanes((size_t)NL);
    uint32_t st = 0;
    bool changed = false;
    int64_t nr = 0;
    AkIdSink sink; sink.buf = nullptr; sink.cap = 0; sink.stride = 1; sink.cnt = 0; sink.direct = true;
    sink.gout = ids; sink.gbase = 0; sink.gcap = id_cap;
    for (int64_t w0 = 0; w0 < n_lanes; w0 += real) {
        for (int l = 0; l < NL; ++l) {
            AkB3Lane& L = lanes[(size_t)l];
            memset(&L, 0, sizeof(L));
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            uint32_t x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 32; ++i) {
                const int64_t q = cs + i;
                if (q >= tb && q < te) { x[i >> 2] |= (uint32_t)text[q] << ((i & 3) * 8); L.own |= 1u << i; }
                if (q >= base0 && q - base0 < (int64_t)rowstart.size() && rowstart[(size_t)(q - base0)]) L.rows |= 1u << i;
            }
            akb3_phase1(x, L);
        }
        for (int l = 0; l < NL; ++l) {
            AkB3Lane& L = lanes[(size_t)l];
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            akb3_phase2(L, l + 1 < NL ? lanes[(size_t)l + 1].dn1 : 0u);
            if (L.FOR) akb3_foreign(T, text, cs, te, L);
            akb3_summary(L);
        }
        for (int l = 0; l < NL; ++l) akb3_phase3(lanes[(size_t)l], l > 0 ? lanes[(size_t)l - 1].up2 : 0u);
        lanes[(size_t)NL - 1].bnd &= 0x3FFFFFFFu;
        for (int l = 1; l <= real; ++l) {
            AkB3Lane& L = lanes[(size_t)l];
            const int64_t cs = base0 + (w0 - 1 + l) * 32;
            const int64_t ss = cs < tb ? tb : cs, se = cs + 32 > te + 1 ? te + 1 : cs + 32;
            if (ss >= se) continue;
            if (L.flags & 1u) st |= AK_ST_ALPHABET;
            for (uint32_t m = L.trb; m;) {
                const int i = akb_ctz(m);
                m &= m - 1u;
                const int64_t p = cs + i;
                const int64_t r = ak_row_lower_bound(off, 0, n_rows, p + 1);
                int64_t cu = -1;
                if (ak_segment_changes(T, text, p, off[r - 1], off[r], 0, &cu, st)) changed = true;
            }
            const uint32_t nb1 = lanes[(size_t)l + 1].bnd, nb2 = l + 2 < NL ? lanes[(size_t)l + 2].bnd : 0u;
            for (uint32_t m = L.rows | L.wstart; m;) {
                const int i = akb_ctz(m);
                m &= m - 1u;
                const int64_t p = cs + i;
                if ((L.rows >> i) & 1u) {
                    while (nr <= n_rows && off[nr] < p) ++nr;
                    while (nr <= n_rows && off[nr] == p) {
                        if (nr > 0 && M.eos >= 0) ak_id_put(sink, M.eos);
                        splits[nr] = sink.cnt;
                        if (nr < n_rows && M.bos >= 0) ak_id_put(sink, M.bos);
                        ++nr;
                    }
                }
                if ((L.wstart >> i) & 1u) {
                    const uint32_t kc = (L.CW >> i) & 1u;
                    const uint32_t above = L.bnd & ~((2u << i) - 1u);
                    int64_t e;
                    if (above) e = cs + akb_ctz(above);
                    else if (nb1) e = cs + 32 + akb_ctz(nb1);
                    else if (nb2) e = cs + 64 + akb_ctz(nb2);
                    else {
                        // the kernel's cold scan (akb3_scan_end)
                        const int64_t er = ak_row_lower_bound(off, 0, n_rows, p + 1);
                        const int64_t re = off[er];
                        int64_t q = cs + (l >= real ? 62 : l == real - 1 ? 94 : 96);
                        if (q > re) q = re;
                        while (q < re && (text[q] & 0xC0u) == 0x80u) ++q;
                        while (q < re) {
                            int len;
                            const uint32_t cp = ak_decode(text, q, re, len);
                            if (AK_HFCLASS(ak_props(T, cp)) != kc) break;
                            q += len;
                        }
                        e = q;
                    }
                    ak_bpe_word(M, T, text, p, e, kc, sink, pool, st);
                }
            }
        }
    }
    *changed_out = changed ? 1 : 0;
    *status = st;
    return sink.cnt;
}

// The fast segment kernel's structure on the CPU (chunks, halo lanes, phase A / B, lane emit, walker slow lane).
void hh_seg_fast(const uint8_t* text, const int64_t* off, int64_t n_rows, uint32_t flags, int real, int stage_cap,
                 int32_t* cluster_ends, int64_t* cluster_splits, int32_t* run_ends, uint8_t* run_tags, int64_t* run_splits,
                 int64_t cap, int64_t* totals, uint32_t* status, int64_t* n_slow) {
    AkTables T = host_tables();
    std::vector<uint32_t> lut(384);
    for (int i = 0; i < 384; ++i) lut[(size_t)i] = i < 128 ? ak_props(T, (uint32_t)i) : ak_props(T, 0x900u + (uint32_t)(i - 128));
    lut.resize(400);
    for (uint32_t ga = 0; ga < 16; ++ga) lut[384 + ga] = ga < 14 ? aks_pair_row(ga) : 0u;
    const bool want_c = (flags & AK_SEG_CLUSTERS) != 0, want_r = (flags & AK_SEG_RUNS) != 0, matras = (flags & AK_SEG_MATRAS) != 0;
    const int64_t tb = off[0], te = off[n_rows], base0 = tb;
    std::vector<uint8_t> rowstart((size_t)(te - base0) + 64, 0);
    for (int64_t r = 0; r <= n_rows; ++r) rowstart[(size_t)(off[r] - base0)] = 1;
    const int64_t n_chunks = (te - base0 + 1 + 15) / 16;
    uint32_t st = 0;
    int64_t cbase = 0, rbase = 0, slow_cnt = 0;
    std::vector<AkSChunk> lanes((size_t)real + 2);
    std::vector<int32_t> cst((size_t)stage_cap + 1), rst((size_t)stage_cap + 1);
    std::vector<uint8_t> tst((size_t)stage_cap + 1);
    for (int64_t w0 = 0; w0 < n_chunks; w0 += real) {
        for (int l = 0; l < real + 2; ++l) {
            AkSChunk& c = lanes[(size_t)l];
            int64_t cs = base0 + (w0 - 1 + l) * 16;
            AkChunk tmp;
            hh_make_chunk(text, cs, tb, te, rowstart, base0, tmp);
            for (int k = 0; k < 5; ++k) c.w[k] = tmp.w[k];
            c.rows = tmp.rows; c.own = tmp.own;
            aks_phase_a(T, lut.data(), c, matras);
        }
        for (int l = 1; l <= real; ++l) {
            AkSChunk c = lanes[(size_t)l];
            int64_t cs = base0 + (w0 - 1 + l) * 16;
            int64_t ss = cs < tb ? tb : cs, se = cs + 16 > te + 1 ? te + 1 : cs + 16;
            if (ss >= se) continue;
            AkSNeighbor pv;
            pv.g = lanes[(size_t)l - 1].end_g; pv.flags = lanes[(size_t)l - 1].flags; pv.end_cur = lanes[(size_t)l - 1].end_cur;
            uint32_t in_cur = AKS_CUR_NONE;
            bool slow = !aks_phase_b(T, lut.data(), c, pv, matras, want_c, want_r, in_cur);
            if (slow) {
                ++slow_cnt;
                AkSegOut o;
                o.cluster_ends = cluster_ends; o.cluster_splits = cluster_splits; o.run_ends = run_ends; o.run_tags = run_tags;
                o.run_splits = run_splits; o.ccap = cap; o.rcap = cap; o.cbase = cbase; o.rbase = rbase;
                int64_t a, b;
                ak_seg_span(T, text, off, n_rows, 0, n_rows, ss, se, flags, 0, true, o, a, b, st);
                cbase += a; rbase += b;
                continue;
            }
            int64_t nr = 0;
            while (nr <= n_rows && off[nr] < ss) ++nr;
            AkSegSink sink;
            sink.cbuf = cst.data(); sink.rbuf = rst.data(); sink.tbuf = tst.data(); sink.cap = stage_cap; sink.stride = 1;
            sink.cc = sink.rc = 0; sink.direct = false; sink.gc = sink.gr = nullptr; sink.gt = nullptr; sink.gccap = sink.grcap = 0;
            int64_t rf, rl;
            aks_lane_emit(c, in_cur, cs, off, n_rows, nr, want_c, want_r, sink, want_c ? cluster_splits : nullptr,
                          want_r ? run_splits : nullptr, rf, rl);
            for (int64_t r = rf; r < rl; ++r) { if (want_c) cluster_splits[r] += cbase; if (want_r) run_splits[r] += rbase; }
            if (sink.cc <= stage_cap && sink.rc <= stage_cap) {
                for (int k = 0; k < sink.cc; ++k) if (cbase + k < cap) cluster_ends[cbase + k] = cst[(size_t)k];
                for (int k = 0; k < sink.rc; ++k) if (rbase + k < cap) { run_ends[rbase + k] = rst[(size_t)k]; run_tags[rbase + k] = tst[(size_t)k]; }
            } else {
                AkSegSink s2 = sink; s2.cc = s2.rc = 0; s2.direct = true; s2.gc = cluster_ends + cbase; s2.gr = run_ends + rbase;
                s2.gt = run_tags + rbase; s2.gccap = cap - cbase; s2.grcap = cap - rbase;
                int64_t a, b;
                aks_lane_emit(c, in_cur, cs, off, n_rows, nr, want_c, want_r, s2, nullptr, nullptr, a, b);
            }
            cbase += sink.cc; rbase += sink.rc;
        }
    }
    totals[0] = cbase; totals[1] = rbase;
    *status = st;
    *n_slow = slow_cnt;
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
