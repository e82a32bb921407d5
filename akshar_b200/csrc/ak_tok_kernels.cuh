// Kernels of the event-stream subword encoders (cores and rationale: ak_tok.cuh).  Included by ak_kernels.cu after the
// common kernel plumbing (ak_common.cuh: AkBatch, ak_batch_begin, akn3_lane_rows, akn3_load_edge).
//
//   ak_words_kernel<KIND>      text -> event slots              (KIND 0: HF Whitespace pre-tokenizer, 1: SentencePiece words)
//   ak_rowfix_kernel           exotic rows -> exact ids in a side pool, their word events struck out
//   ak_resolve_kernel<KIND>    event slot -> resolved record (word cache look-up, exact encoder on a miss)
//   ak_unicheck_kernel         Unigram: which cached word lattices need the exact Viterbi of their row
//   ak_emit_kernel             resolved records -> ids + row splits at their final place
#pragma once

#define AKT_WARP_BYTES 960                             // text bytes per warp tile (30 lanes x 32 bytes)
#define AKW_THREADS 128                                // words kernel: four independent warps
#define AKR_THREADS 256                                // resolve kernel
#define AKL_THREADS 256                                // check / emit kernels
#define AKL_PER 4                                      // slots per thread
#define AKL_TILE (AKL_THREADS * AKL_PER)               // 1024 slots per CTA tile
#define AKT_LONG_ROW 8192                              // Unigram: longer rows go to the exact row encoder

// event slots: warp tile w owns slots [w * cap, (w + 1) * cap), the first count[w] of them hold its events in text order
struct AkSlots {
    AkEvent* ev;
    uint32_t* count;
    int cap;                           // a power of two >= 256: slot -> warp tile is a shift
    int shift;                         // log2(cap)
};

__device__ __forceinline__ long long akt_n_wt(const AkBatch& B, int64_t base0) {
    return (B.text_end - base0 + AKT_WARP_BYTES) / AKT_WARP_BYTES;      // covers position text_end itself
}

// The word cache lives as long as its model (like HF's): words learned in one call serve the next.  It has no eviction,
// so once the entries added since the last restore pass `limit` the pristine image (built at model load) is copied back
// over it at the start of a call -- decided on the device, no host synchronisation.  force != 0 restores unconditionally.
__global__ void ak_cache_guard_kernel(unsigned long long* work, const unsigned long long* image, size_t n_words,
                                      const unsigned long long* inserted, unsigned long long limit, int force) {
    if (!force && *inserted <= limit) return;
    const size_t n2 = n_words / 2;
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(image);
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(work);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void ak_cache_guard_reset_kernel(unsigned long long* inserted, unsigned long long limit, int force) {
    if (force || *inserted > limit) *inserted = 0ull;
}

// =================================================================================================================
// words kernel
// =================================================================================================================
struct AkWordsArgs {
    AkBatch B;
    AkTables T;
    const int64_t* wrow;               // first row at or after base0 + 480 k
    int64_t base0;
    AkSlots S;
    uint8_t* row_flag;                 // per row: 1 = exotic (encoded by the row-fix kernel)
    unsigned int* any_flag;
    uint32_t* row_ev;                  // [n_rows + 1] slot of the event that starts the row
    long long* n_wt_out;               // number of warp tiles (for the scans over the per-warp-tile aggregates)
    AkBpeDev bpe;                      // KIND 0: the added tokens
};

// end of a word with no boundary within what the warp knows (cold)
template <int KIND>
struct AkScanEnd {
    const AkTables* T;
    const uint8_t* text;
    const int64_t* off;
    int64_t n_rows, cs;
    uint32_t cw;
    __device__ int64_t operator()(int64_t p, int64_t from) const {
        if (KIND == 0) return akb3_scan_end(*T, text, p, from, (cw >> (int)(p - cs)) & 1u, off, n_rows, 0, n_rows);
        // SentencePiece word: up to the next space or the end of the row
        const int64_t er = ak_row_lower_bound(off, 0, n_rows, p + 1);
        const int64_t re = off[er];
        int64_t q = from > re ? re : from;
        while (q < re && text[q] != 0x20u) ++q;
        return q;
    }
};

#ifndef AKW_MINB
#define AKW_MINB 8
#endif
template <int KIND>
__global__ void __launch_bounds__(AKW_THREADS, AKW_MINB) ak_words_kernel(const AkWordsArgs A) {
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tb = B.text_begin, te = B.text_end;
    const long long n_wt = akt_n_wt(B, A.base0);
    if (blockIdx.x == 0 && threadIdx.x == 0) *A.n_wt_out = n_wt;
    uint32_t st = 0;
    for (long long wt = (long long)blockIdx.x * (AKW_THREADS / 32) + warp; wt < n_wt; wt += (long long)gridDim.x * (AKW_THREADS / 32)) {
        const int64_t ws0 = A.base0 + wt * AKT_WARP_BYTES;
        const int64_t cs = ws0 + (int64_t)(lane - 1) * 32;
        const int64_t r_w0 = A.wrow[2 * wt];
        uint32_t x[8];
        uint32_t own;
        {
            int64_t lo = tb - cs, hi = te - cs;
            lo = lo < 0 ? 0 : (lo > 32 ? 32 : lo);
            hi = hi < 0 ? 0 : (hi > 32 ? 32 : hi);
            if (lo == 0 && hi == 32) {
                const uint4 v0 = *reinterpret_cast<const uint4*>(B.text + cs);
                const uint4 v1 = *reinterpret_cast<const uint4*>(B.text + cs + 16);
                x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
                x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                own = 0xFFFFFFFFu;
            } else {
                akn3_load_edge(B.text, cs, (int)lo, (int)hi, x);
                own = hi > lo ? ((hi == 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u)) : 0u;
            }
        }
        int64_t first_row;
        int nrows;
        const uint32_t rows = akn3_lane_rows2(B.off, B.n_rows, r_w0, ws0, lane, first_row, nrows);
        const bool real = lane >= 1 && lane <= 30;
        const int64_t ss = cs < tb ? tb : cs;
        const int64_t se = cs + 32 > te + 1 ? te + 1 : cs + 32;
        const bool active = real && ss < se;
        uint32_t rowsm, wstart, cw, bnd, nb1, nb2;
        if (KIND == 0) {
            AkB3Lane L;
            L.own = own;
            L.rows = rows;
            akb3_phase1(x, L);
            uint32_t dn1n = __shfl_down_sync(0xFFFFFFFFu, L.dn1, 1);
            if (lane == 31) dn1n = 0;
            akb3_phase2(L, dn1n);
            if (L.FOR) akb3_foreign(A.T, B.text, cs, te, L);
            akb3_summary(L);
            const uint32_t up2p = __shfl_up_sync(0xFFFFFFFFu, L.up2, 1);
            akb3_phase3(L, up2p);
            if (active) {
                // what only HF's side of the pipeline acts on: an added token in the raw text ('<' is where each of the
                // reference's five starts), a code point its NFKC changes -> the row is encoded by the row-fix kernel
                uint32_t fix = L.UNS & L.own;
                for (uint32_t m = L.LT & L.own; m;) {
                    const int i = akb_ctz(m);
                    m &= m - 1u;
                    if (ak_bpe_special_at(A.bpe, B.text, cs + i, te) >= 0) fix |= 1u << i;
                }
                if (fix) {
                    ake_flag_rows(B.off, B.n_rows, cs, fix, A.row_flag);
                    atomicOr(A.any_flag, 1u);
                }
                if (L.trb) {
                    const int64_t r_lo = r_w0 > 0 ? r_w0 - 1 : 0;
                    if (akb3_changes(A.T, B.text, B.off, B.n_rows, r_lo, L.trb, cs, st)) {
                        // NFC (which HF's NFKC includes) would change a code point of this lane: its row is normalized
                        // and encoded on its own by the row-fix kernel
                        ake_flag_rows(B.off, B.n_rows, cs, L.trb, A.row_flag);
                        atomicOr(A.any_flag, 1u);
                    }
                }
            }
            const uint32_t bsend = lane == 31 ? (L.bnd & 0x3FFFFFFFu) : L.bnd;
            nb1 = __shfl_down_sync(0xFFFFFFFFu, bsend, 1);
            nb2 = __shfl_down_sync(0xFFFFFFFFu, bsend, 2);
            if (lane >= 30) nb2 = 0;
            wstart = active ? L.wstart : 0u;
            rowsm = active ? L.rows : 0u;
            cw = L.CW;
            bnd = L.bnd;
        } else {
            AkU3Lane L;
            L.own = own;
            L.rows = rows;
            aku3_phase1(x, L);
            uint32_t upp = __shfl_up_sync(0xFFFFFFFFu, L.up, 1);
            if (lane == 0) upp = 0;
            aku3_phase2(L, upp);
            nb1 = __shfl_down_sync(0xFFFFFFFFu, L.bnd, 1);
            nb2 = __shfl_down_sync(0xFFFFFFFFu, L.bnd, 2);
            if (lane >= 30) nb2 = 0;
            wstart = active ? L.wstart : 0u;
            rowsm = active ? L.rows : 0u;
            cw = 0xFFFFFFFFu;
            bnd = L.bnd;
            if (active && L.exotic) {
                // a (possible) literal U+2581: the rows that hold it go to the exact row encoder
                ake_flag_rows(B.off, B.n_rows, cs, L.exotic, A.row_flag);
                atomicOr(A.any_flag, 1u);
            }
        }
        // the lane's place among the warp tile's slots
        const int n_ev = __popc(rowsm) + __popc(wstart);
        int inc = n_ev;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += y;
        }
        const int total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        const int pre = inc - n_ev;
        if (lane == 0) {
            A.S.count[wt] = (uint32_t)(total < A.S.cap ? total : A.S.cap);
            if (total > A.S.cap) {
                // too many events in 960 bytes for the slots this workspace gives a warp tile: say how many were needed
                st |= AK_ST_OVERFLOW;
                atomicMax((unsigned long long*)&B.result[3], (unsigned long long)total);
            }
        }
        if (n_ev) {
            const int64_t at = (wt << A.S.shift) + pre;
            const int lane_tail = lane >= 30 ? 62 : lane == 29 ? 94 : 96;       // bytes from cs on whose boundaries the warp knows
            AkScanEnd<KIND> se;
            se.T = &A.T; se.text = B.text; se.off = B.off; se.n_rows = B.n_rows; se.cs = cs; se.cw = cw;
            ake_lane_events(rowsm, wstart, cw, bnd, nb1, nb2, lane_tail, cs, tb, B.off, B.n_rows, first_row, nrows, at, A.row_ev,
                            A.S.ev + at, (int64_t)A.S.cap - pre, se);
        }
    }
    ak_raise(B.result, st);
}

// Unigram: rows longer than AKT_LONG_ROW go to the exact row encoder (the word-wise path would send most of their words to
// the exact Viterbi anyway: the accumulated score grows with the row)
__global__ void ak_long_rows_kernel(AkBatch B, int64_t limit, uint8_t* row_flag, unsigned int* any_flag) {
    if (!ak_batch_begin(B)) return;
    bool any = false;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < B.n_rows; g += (int64_t)gridDim.x * blockDim.x)
        if (B.off[g + 1] - B.off[g] > limit) { row_flag[g] = 1; any = true; }
    if (any) atomicOr(any_flag, 1u);
}

// =================================================================================================================
// row fix: the rows the fast path does not handle (BPE: not in NFC; Unigram: a literal U+2581 or a row longer than
// AKT_LONG_ROW) are encoded whole by the exact row encoder into a side pool; their word events are struck out and the
// row's start event hands the ids over.  One thread per flagged row (rare).
// =================================================================================================================
struct AkRowFixArgs {
    AkBatch B;
    AkRowFixCtx X;
    int64_t base0;
    int cap;
    const uint8_t* row_flag;
    const unsigned int* any_flag;
};

__global__ void __launch_bounds__(128) ak_rowfix_kernel(const AkRowFixArgs A) {
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    if (*A.any_flag == 0u) return;
    AkRowFixCtx X = A.X;
    X.n_events = (unsigned long long)akt_n_wt(B, A.base0) * (unsigned long long)A.cap;
    X.text = B.text;
    X.off = B.off;
    X.n_rows = B.n_rows;
    X.result = B.result;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < B.n_rows; g += (int64_t)gridDim.x * blockDim.x)
        if (A.row_flag[g]) akr_fix_row(X, g);
}

// =================================================================================================================
// resolve kernel: one WARP per warp tile, one slot per lane and round.  The lanes of a round run the same look-up side by
// side and meet again (__syncwarp) before the next one -- without that, lanes that finish early (an empty slot, a short
// word) run ahead through the loop on their own and the warp degenerates into 32 single-lane instruction streams.
// Per warp tile it leaves the number of ids its events emit and (Unigram) the segmented sum of wmag.
// =================================================================================================================
struct AkResolveArgs {
    AkBatch B;
    AkLookupCtx X;
    int64_t base0;
    AkSlots S;
    unsigned long long* resolved;      // per slot
    uint32_t* aux;                     // per slot (Unigram)
    int32_t* wt_ids;                   // per warp tile: ids emitted
    unsigned long long* wt_seg;        // per warp tile (Unigram): (has a row start << 32) | float sum of wmag after the last one
    const unsigned int* any_flag;
};

// ---- Unigram: the lattice of one word, solved by the whole warp --------------------------------------------------------
// A word that is not in the cache needs its lattice (aku_word_lattice: a trie walk from every code point, then the Viterbi
// recursion).  One lane doing that alone holds up the 31 others of its round for thousands of dependent instructions; here
// the warp does it together: the lanes decode the word, every lane walks the trie from its own start position (the longest
// dependent chain is now one walk, not the sum of all), and the recursion takes the candidates of a position from 17 lanes
// at once.  Same arithmetic (double sums of the float scores, first-come tie break = longest piece), same ratio / wmag.
#define AKU_W_KMAX 16                                  // pieces of up to 16 code points (the trainer's max_sentencepiece_length)
struct AkUniWarpScratch {
    uint32_t cps[AKU_WORD_CPS];
    uint32_t epid[AKU_WORD_CPS][AKU_W_KMAX];           // [start][length - 1] = piece id + 1 of a usable piece, else 0; later: the ids
    double best[AKU_WORD_CPS + 1];
    uint32_t bk[AKU_WORD_CPS + 1];
    int n_ids;
    float ratio, wmag;
};

// warp maximum of a double (no NaNs): the bits mapped to an unsigned key of the same order, one `redux.sync` for the high
// words, one for the low words of the lanes that hold the highest -- a selection, so the result is the exact maximum
__device__ __forceinline__ double aku_warp_max_d(double v) {
    const uint32_t hi = (uint32_t)__double2hiint(v), lo = (uint32_t)__double2loint(v);
    const uint32_t neg = (uint32_t)((int32_t)hi >> 31);
    const uint32_t khi = hi ^ (neg | 0x80000000u), klo = lo ^ neg;
    const uint32_t H = __reduce_max_sync(0xFFFFFFFFu, khi);
    const uint32_t L = __reduce_max_sync(0xFFFFFFFFu, khi == H ? klo : 0u);
    const uint32_t back = (H & 0x80000000u) ? 0u : 0xFFFFFFFFu;
    return __hiloint2double((int)(H ^ (back | 0x80000000u)), (int)(L ^ back));
}

// all 32 lanes call this with the same word; false = the word does not fit the cooperative scheme (the caller's serial path
// takes it).  The result is in S (ids in S.epid as int32).
__device__ __noinline__ bool aku_word_lattice_warp(const AkUniDev& U, const uint8_t* t, int64_t s, uint32_t len, AkUniWarpScratch& S) {
    const int lane = threadIdx.x & 31;
    if (len > 2u * 32u - 8u) return false;
    // code points: lane l looks at bytes 2l and 2l + 1
    uint32_t mine = 0;
    uint32_t c[2] = {0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t q = 2u * (uint32_t)lane + (uint32_t)h;
        if (q < len && (t[s + q] & 0xC0u) != 0x80u) {
            int l;
            c[h] = ak_decode(t, s + q, s + len, l);
            mine |= 1u << h;
        }
    }
    int inc = __popc(mine);
    const int own = inc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += y;
    }
    const int n = 1 + __shfl_sync(0xFFFFFFFFu, inc, 31);
    if (n > AKU_WORD_CPS) return false;
    {
        int at = 1 + inc - own;
        if (mine & 1u) S.cps[at++] = c[0];
        if (mine & 2u) S.cps[at] = c[1];
        if (lane == 0) S.cps[0] = (U.flags & 4) ? 0x2581u : 0x20u;
    }
    __syncwarp();
    // trie walks: start positions lane and lane + 32
    bool deep = false;
    for (int i = lane; i < n; i += 32) {
#pragma unroll
        for (int k = 0; k < AKU_W_KMAX; ++k) S.epid[i][k] = 0u;
        uint32_t node = 0;
        for (int k = 1; i + k <= n; ++k) {
            const unsigned long long v = ak_uni_child(U, node, S.cps[i + k - 1]);
            if (v == AK_EMPTY_KEY) break;
            if (k > AKU_W_KMAX) { deep = true; break; }
            node = (uint32_t)(v >> 32);
            const uint32_t pid1 = (uint32_t)v;
            if (pid1 && U.usable[pid1 - 1u]) S.epid[i][k - 1] = pid1;
        }
    }
    if (__any_sync(0xFFFFFFFFu, deep)) return false;
    if (lane == 0) { S.best[0] = 0.0; S.bk[0] = 0u; }
    __syncwarp();
    // the recursion, position by position: lanes 1 .. 16 bring the pieces that end here, lane 0 the unknown-character edge
    double margin = INFINITY, wmag = 0.0;
    for (int j = 1; j <= n; ++j) {
        double cand = -INFINITY;
        uint32_t b = 0u;
        if (lane >= 1 && lane <= AKU_W_KMAX && j - lane >= 0) {
            const int i = j - lane;
            const uint32_t pid1 = S.epid[i][lane - 1];
            const double bi = S.best[i];
            if (pid1 && bi != -INFINITY) {
                cand = bi + (double)U.score[pid1 - 1u];
                b = ((uint32_t)lane << 24) | (pid1 - 1u);
            }
        } else if (lane == 0) {
            const double bi = S.best[j - 1];
            if (bi != -INFINITY && S.epid[j - 1][0] == 0u) {
                cand = bi + (double)U.unk_score;
                b = AK_UNI_UNKBIT | S.cps[j - 1];
            }
        }
        const double a = cand == -INFINITY ? 0.0 : (cand < 0 ? -cand : cand);
        const double m1 = aku_warp_max_d(cand);
        const double wm = aku_warp_max_d(a);
        if (wm > wmag) wmag = wm;
        // the first candidate to arrive with the best value wins: the longest piece (the unknown edge comes last)
        const unsigned tie = __ballot_sync(0xFFFFFFFFu, cand == m1 && m1 != -INFINITY);
        const int win = tie ? 31 - __clz((int)tie) : 0;
        const double m2 = aku_warp_max_d(lane == win ? -INFINITY : cand);
        const uint32_t wb = __shfl_sync(0xFFFFFFFFu, b, win);
        if (m1 != -INFINITY && m1 - m2 < margin) margin = m1 - m2;
        if (lane == 0) { S.best[j] = m1; S.bk[j] = wb; }
        __syncwarp();
    }
    // ids, back to front (one lane; the edge table is done with and takes them)
    if (lane == 0) {
        int32_t* ids = reinterpret_cast<int32_t*>(&S.epid[0][0]);
        int cnt = 0;
        for (int j = n; j > 0;) {
            const uint32_t b = S.bk[j];
            if (b & AK_UNI_UNKBIT) { cnt += (U.flags & 8) ? ak_utf8_len(b & 0x1FFFFFu) : 1; j -= 1; }
            else { cnt += 1; j -= (int)((b >> 24) & 0x7Fu); }
        }
        S.n_ids = cnt;
        int at = cnt;
        for (int j = n; j > 0;) {
            const uint32_t b = S.bk[j];
            if (b & AK_UNI_UNKBIT) {
                const uint32_t cp = b & 0x1FFFFFu;
                if (U.flags & 8) {
                    uint8_t enc[4];
                    const int m = ak_encode(cp, enc);
                    for (int q = m - 1; q >= 0; --q) ids[--at] = U.byte_id[enc[q]];
                } else ids[--at] = U.unk_id;
                j -= 1;
            } else {
                ids[--at] = (int32_t)(b & 0xFFFFFFu);
                j -= (int)((b >> 24) & 0x7Fu);
            }
        }
        const double r = margin / (double)n;
        S.ratio = r > 1e30 ? 1e30f : (float)r * 0.999f;
        S.wmag = (float)wmag * 1.001f + 1e-30f;
    }
    __syncwarp();
    return true;
}

// CTAs per SM the resolve kernel is compiled for: 4 (64 registers) since the next round's event is fetched ahead -- the loop
// is bound by round trips, and a fourth CTA's warps hide more of them than the 14 extra registers of 3 CTAs saved
// (3.38 -> 2.83 ms BPE, 4.07 -> 3.76 ms Unigram; 5 and 6 CTAs: 2.89 / 2.85 ms BPE, 4.46 ms Unigram)
#ifndef AKR_MINB0
#define AKR_MINB0 4
#endif
#ifndef AKR_MINB1
#define AKR_MINB1 4
#endif
#ifndef AKE_MINB
#define AKE_MINB 4                                     // emit, with the rare records deferred per four-round block: 4 / 5 / 6 / 3 CTAs per SM = 0.92 / 1.01 / 1.17 / 1.12 ms (64 registers at 4)
#endif
#ifndef AKC_MINB
#define AKC_MINB 8                                     // unicheck: latency bound, 32 registers are enough (12.15 -> 11.56 ms per GiB step)
#endif
template <int KIND>
__global__ void __launch_bounds__(AKR_THREADS, KIND == 0 ? AKR_MINB0 : AKR_MINB1) ak_resolve_kernel(const AkResolveArgs A) {
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    AkLookupCtx X = A.X;
    X.text = B.text;
    X.off = B.off;
    X.n_rows = B.n_rows;
    X.tb = B.text_begin;
    X.te = B.text_end;
    X.result = B.result;
    X.any_fix = *A.any_flag != 0u ? 1 : 0;
    const int lane = threadIdx.x & 31;
    __shared__ AkUniWarpScratch s_uni[KIND == 1 ? AKR_THREADS / 32 : 1];
    AkUniWarpScratch* const scratch = &s_uni[KIND == 1 ? (threadIdx.x >> 5) : 0];
    const long long n_wt = akt_n_wt(B, A.base0);
    const long long warp0 = ((long long)blockIdx.x * AKR_THREADS + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * AKR_THREADS) >> 5;
    uint32_t st = 0;
    for (long long wt = warp0; wt < n_wt; wt += n_warps) {
        const int cnt = (int)A.S.count[wt];
        const long long s_wt = wt << A.S.shift;
        int ids = 0;
        unsigned long long seg = 0ull;                         // (flag, sum) over the rounds so far
        // the next round's event is fetched while this round is looked up (one round trip less in the chain)
        AkEvent ev_next;
        ev_next.pos = 0u;
        ev_next.meta = AKE_DEAD;
        if (lane < cnt) ev_next = A.S.ev[s_wt + lane];
        for (int o = 0; o < cnt; o += 32) {
            const int i = o + lane;
            const AkEvent ev_cur = ev_next;
            if (i + 32 < cnt) ev_next = A.S.ev[s_wt + i + 32];
            unsigned long long r = 0ull;
            uint32_t aux = 0u;
            bool miss = false;
            int64_t miss_p = 0;
            uint32_t miss_len = 0u;
            AkcHit hit;
            hit.slot = hit.free_slot = -1;
            hit.h = hit.want = hit.tag = hit.ids01 = 0ull;
            if (i < cnt) {
                AkEvent ev = ev_cur;
                const uint32_t kind = ev.meta & 7u, len = ev.meta >> 3;
                unsigned long long k[4] = {0ull, 0ull, 0ull, 0ull};
                if (kind <= AKE_WORD && len <= AKC_MAXLEN) akc_key0123(X.text, X.tb + ev.pos, len, X.te, k);
                if (KIND == 1 && kind <= AKE_WORD && len <= 56u && len != AKE_LEN_MAX) {
                    // Unigram word: look it up here; a miss is solved by the whole warp below
                    akc_lookup4(X.M.cache, X.text, X.tb + ev.pos, len, k, hit);
                    if (hit.slot >= 0) {
                        const int n = AKC_NTOK(hit.tag);
                        aux = (uint32_t)(hit.tag >> 32);
                        r = n <= 2 ? akr_inline(n, hit.ids01) : akr_cache(n, hit.slot);
                    } else {
                        miss = true;
                        miss_p = X.tb + ev.pos;
                        miss_len = len;
                    }
                } else r = akl_resolve<KIND>(X, ev, k, aux, st);
            }
            if (KIND == 1) {
                for (unsigned mm = __ballot_sync(0xFFFFFFFFu, miss); mm; mm &= mm - 1u) {
                    const int src = __ffs((int)mm) - 1;
                    const int64_t wp = __shfl_sync(0xFFFFFFFFu, miss_p, src);
                    const uint32_t wl = __shfl_sync(0xFFFFFFFFu, miss_len, src);
                    const bool solved = aku_word_lattice_warp(X.M.uni, X.text, wp, wl, *scratch);
                    if (lane == src) {
                        AkMissOut o;
                        if (solved)
                            o = akl_uni_finish(X, wp, wl, hit.free_slot, hit.want, true, true, scratch->n_ids,
                                               reinterpret_cast<const int32_t*>(&scratch->epid[0][0]), scratch->ratio, scratch->wmag);
                        else o = akl_uni_miss(X, wp, wl, hit.free_slot, hit.want, true);
                        st |= o.st;
                        aux = (uint32_t)(akc_aux(o.ratio, o.wmag) >> 32);
                        r = o.slot == -1 ? akr_inline(o.n, o.ids01) : akr_pool(o.n, (unsigned long long)(-3 - o.slot));
                    }
                    __syncwarp();
                }
            }
            if (i < cnt) {
                A.resolved[s_wt + i] = r;
                if (KIND == 1) A.aux[s_wt + i] = aux;
            }
            __syncwarp();
            ids += akr_n(r);
            if (KIND == 1) {
                // my element, then the warp's aggregate of this round appended to the running one
                unsigned long long e = (r >> 62) == AKR_EVENT ? AKS_SEG_FLAG : (unsigned long long)__float_as_uint(aku_aux_wmag(aux));
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, e, d);
                    if (lane >= d) e = aks_seg_op(y, e);
                }
                seg = aks_seg_op(seg, __shfl_sync(0xFFFFFFFFu, e, 31));
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ids += __shfl_xor_sync(0xFFFFFFFFu, ids, d);
        if (lane == 0) {
            A.wt_ids[wt] = ids;
            if (KIND == 1) A.wt_seg[wt] = seg;
        }
    }
    ak_raise(B.result, st);
}

// =================================================================================================================
// Unigram check: d = segmented sum of wmag along each row (a row event starts a segment); a word whose cached lattice
// is not robust at that magnitude takes its ids from the exact Viterbi of its row.  One warp per warp tile; the sum that
// reaches the warp tile comes from the scan over the warp tiles' aggregates (ak_scan_seg_kernel).
// =================================================================================================================
struct AkCheckArgs {
    AkBatch B;
    AkLookupCtx X;
    int64_t base0;
    AkSlots S;
    unsigned long long* resolved;
    const uint32_t* aux;
    const float* wt_seg_before;        // per warp tile: sum of wmag since the last row start before it
    int32_t* wt_ids;
    const unsigned int* any_flag;
};

__global__ void __launch_bounds__(AKL_THREADS, AKC_MINB) ak_unicheck_kernel(const AkCheckArgs A) {
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    AkLookupCtx X = A.X;
    X.text = B.text;
    X.off = B.off;
    X.n_rows = B.n_rows;
    X.tb = B.text_begin;
    X.te = B.text_end;
    X.result = B.result;
    X.any_fix = *A.any_flag != 0u ? 1 : 0;
    const int lane = threadIdx.x & 31;
    const long long n_wt = akt_n_wt(B, A.base0);
    const long long warp0 = ((long long)blockIdx.x * AKL_THREADS + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * AKL_THREADS) >> 5;
    for (long long wt = warp0; wt < n_wt; wt += n_warps) {
        const int cnt = (int)A.S.count[wt];
        const long long s_wt = wt << A.S.shift;
        unsigned long long run = (unsigned long long)__float_as_uint(A.wt_seg_before[wt]);     // (flag, sum) before this round
        int delta = 0;
        for (int o4 = 0; o4 < cnt; o4 += 128) {
          // the loads of four rounds go out together (a round's work is a dependent chain of shuffles behind them)
          uint32_t aux4[4];
          bool row4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
              const int i = o4 + 32 * q + lane;
              aux4[q] = 0u;
              row4[q] = false;
              if (i < cnt) {
                  aux4[q] = A.aux[s_wt + i];
                  row4[q] = (A.resolved[s_wt + i] >> 62) == AKR_EVENT;
              }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int o = o4 + 32 * q;
            if (o >= cnt) break;
            const int i = o + lane;
            const uint32_t aux = aux4[q];
            const bool isrow = row4[q];
            unsigned long long e = isrow ? AKS_SEG_FLAG : (unsigned long long)__float_as_uint(aku_aux_wmag(aux));
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, e, d);
                if (lane >= d) e = aks_seg_op(y, e);
            }
            const unsigned long long incl = aks_seg_op(run, e);                 // up to and including my slot
            run = aks_seg_op(run, __shfl_sync(0xFFFFFFFFu, e, 31));
            if (i < cnt && !isrow && aux != 0u && !aku_robust(aku_aux_ratio(aux), __uint_as_float((uint32_t)incl))) {
                const AkEvent ev = A.S.ev[s_wt + i];
                if ((ev.meta & 7u) <= AKE_WORD) {
                    long long pa;
                    const int n = akl_uni_exact(X, X.tb + ev.pos, ev.meta >> 3, &pa);
                    const unsigned long long r = pa >= 0 ? akr_pool(n, (unsigned long long)pa) : 0ull;
                    delta += akr_n(r) - akr_n(A.resolved[s_wt + i]);
                    A.resolved[s_wt + i] = r;
                }
            }
            __syncwarp();
          }
        }
        if (delta) atomicAdd(&A.wt_ids[wt], delta);
    }
}

// =================================================================================================================
// emit kernel: one warp per warp tile; its first id goes to wt_base[wt] (scan over the warp tiles' id counts)
// =================================================================================================================
struct AkEmitArgs {
    AkBatch B;
    AkLookupCtx X;
    int64_t base0;
    AkSlots S;
    const unsigned long long* resolved;
    const int64_t* wt_base;
    const unsigned int* any_flag;
};

template <class IdT>
__global__ void __launch_bounds__(AKL_THREADS, AKE_MINB) ak_emit_kernel(const AkEmitArgs A) {
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    AkLookupCtx X = A.X;
    X.text = B.text;
    X.off = B.off;
    X.n_rows = B.n_rows;
    X.tb = B.text_begin;
    X.te = B.text_end;
    X.result = B.result;
    X.any_fix = *A.any_flag != 0u ? 1 : 0;
    IdT* const ids = (IdT*)A.X.ids;
    const int64_t id_cap = A.X.id_cap;
    const int32_t bos = A.X.M.kind == 0 ? A.X.M.bpe.bos : -1, eos = A.X.M.kind == 0 ? A.X.M.bpe.eos : -1;
    const int splits_i32 = A.X.splits_i32;
    const unsigned long long* const cache_e = A.X.M.cache.e;
    const unsigned long long cache_entries = 1ull << A.X.M.cache.bits;
    const int lane = threadIdx.x & 31;
    const long long n_wt = akt_n_wt(B, A.base0);
    const long long warp0 = ((long long)blockIdx.x * AKL_THREADS + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * AKL_THREADS) >> 5;
    uint32_t st = 0;
    for (long long wt = warp0; wt < n_wt; wt += n_warps) {
        const int cnt = (int)A.S.count[wt];
        const long long s_wt = wt << A.S.shift;
        int64_t at0 = A.wt_base[wt];
        for (int o4 = 0; o4 < cnt; o4 += 128) {
          unsigned long long r4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
              const int i = o4 + 32 * q + lane;
              r4[q] = i < cnt ? A.resolved[s_wt + i] : 0ull;
          }
          // the four rounds: the scans, and the common case (one or two ids held in the record itself) at once.  Row events
          // and cache-entry records are rare per lane but present in nearly every round: they wait (their round in `later`,
          // their place in rel4) and are written after the four scans, each lane walking its own -- ~8 lanes busy per
          // pass instead of ~2 in every round.
          const int64_t at_blk = at0;
          int rel4[4] = {0, 0, 0, 0};
          uint32_t later = 0u;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int o = o4 + 32 * q;
            if (o >= cnt) break;
            const unsigned long long r = r4[q];
            const unsigned long long ty = r >> 62;
            const int n = ty == AKR_INLINE ? (int)((r >> 60) & 3ull) : akr_n(r);
            int inc = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += y;
            }
            const int64_t at = at0 + (inc - n);
            rel4[q] = (int)(at - at_blk);
            if (ty == AKR_INLINE) {
                if (n > 0) {
                    if (at + n > id_cap) st |= AK_ST_OVERFLOW;
                    else {
                        ids[at] = (IdT)(r & 0x3FFFFFFFull);
                        if (n > 1) ids[at + 1] = (IdT)((r >> 30) & 0x3FFFFFFFull);
                    }
                }
            } else if (o + lane < cnt) later |= 1u << q;
            at0 += __shfl_sync(0xFFFFFFFFu, inc, 31);
          }
          while (later) {
            const int q = __ffs(later) - 1;
            later &= later - 1u;
            const unsigned long long r = q == 0 ? r4[0] : q == 1 ? r4[1] : q == 2 ? r4[2] : r4[3];
            const int64_t at = at_blk + (q == 0 ? rel4[0] : q == 1 ? rel4[1] : q == 2 ? rel4[2] : rel4[3]);
            const int i = o4 + 32 * q + lane;
            const unsigned long long ty = r >> 62;
            const int n = akr_n(r);
            if (ty == AKR_EVENT && !(r & (1ull << 37))) {
                // one (unfixed) row starts here: </s> of the previous row, the split, <s>
                const int64_t g = (int64_t)(r & ((1ull << 37) - 1ull));
                int64_t k2 = at;
                if (at + n > id_cap) st |= AK_ST_OVERFLOW;
                else {
                    if (g > 0 && eos >= 0) ids[k2++] = (IdT)eos;
                    if (splits_i32) ((int32_t*)A.X.splits)[g] = (int32_t)k2;
                    else ((int64_t*)A.X.splits)[g] = k2;
                    if (g < B.n_rows && bos >= 0) ids[k2] = (IdT)bos;
                }
            } else if (ty == AKR_CACHE && (r & 0xFFFFFFFFFFFFFFull) < cache_entries && n <= AKC_MAXTOK) {
                // three to fourteen ids held in the word's cache entry: ids 0-1 in word 3, the rest two per word from word 9
                if (at + n > id_cap) st |= AK_ST_OVERFLOW;
                else {
                    const unsigned long long* en = cache_e + (r & 0xFFFFFFFFFFFFFFull) * AKC_ENTRY;
                    unsigned long long v = akc_ld(en + 3);
                    ids[at] = (IdT)(uint32_t)v;
                    ids[at + 1] = (IdT)(uint32_t)(v >> 32);
                    for (int j = 2; j < n; j += 2) {
                        v = akc_ld(en + 8 + (j >> 1));
                        ids[at + j] = (IdT)(uint32_t)v;
                        if (j + 1 < n) ids[at + j + 1] = (IdT)(uint32_t)(v >> 32);
                    }
                }
            } else {
                if (at + n > id_cap) st |= AK_ST_OVERFLOW;
                akl_emit(X, r, A.S.ev + s_wt + i, at);
            }
          }
          __syncwarp();
        }
    }
    ak_raise(B.result, st);
}
