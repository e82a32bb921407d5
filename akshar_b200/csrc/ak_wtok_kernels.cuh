// Kernels of the word tokenizers (cores: ak_wordtok.cuh).  Two passes over the text, both warp-autonomous (a warp owns
// 960 text bytes: 30 real lanes + 2 halo lanes): the count pass leaves the number of tokens that start in every warp
// tile, one scan (ak_scan_counts_kernel) turns them into bases, the emit pass classifies again and writes every token's
// begin / end and every row's split at its final place -- no temporary stream, no copy kernel, no ordering between warps.
// Algorithmic bytes per text byte: 2 read (the text, twice) + 8 per token written (+ 8 per row).
#pragma once
#include "ak_wordtok.cuh"

#define AKWT_THREADS 128

struct AkWtArgs {
    AkBatch B;
    const int64_t* wrow;               // first row at or after base0 + 480 k
    int64_t base0;
    int mode;
    int32_t* count;                    // [n_wt] tokens that start in the warp tile
    const int64_t* base;               // [n_wt] exclusive prefix of count
    int32_t* begin;                    // [cap] byte offset of the token's first byte, relative to its row
    int32_t* end;                      // [cap] ... of the byte after its last
    int64_t cap;
    int64_t* splits;                   // [n_rows + 1]
    uint8_t* row_flags;                // optional [n_rows]: bit 0 = the row holds a code point of U+0900-097F
};

template <bool EMIT>
__global__ void __launch_bounds__(AKWT_THREADS, 8) ak_wtok_kernel(const AkWtArgs A) {
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tb = B.text_begin, te = B.text_end;
    const long long n_wt = akt_n_wt(B, A.base0);
    uint32_t st = 0;
    for (long long wt = (long long)blockIdx.x * (AKWT_THREADS / 32) + warp; wt < n_wt; wt += (long long)gridDim.x * (AKWT_THREADS / 32)) {
        const int64_t ws0 = A.base0 + wt * AKT_WARP_BYTES;
        const int64_t cs = ws0 + (int64_t)(lane - 1) * 32;
        const int64_t r_w0 = A.wrow[2 * wt];
        AkWtLane L;
        uint32_t x[8];
        {
            int64_t lo = tb - cs, hi = te - cs;
            lo = lo < 0 ? 0 : (lo > 32 ? 32 : lo);
            hi = hi < 0 ? 0 : (hi > 32 ? 32 : hi);
            if (lo == 0 && hi == 32) {
                const uint4 v0 = *reinterpret_cast<const uint4*>(B.text + cs);
                const uint4 v1 = *reinterpret_cast<const uint4*>(B.text + cs + 16);
                x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
                x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                L.own = 0xFFFFFFFFu;
            } else {
                akn3_load_edge(B.text, cs, (int)lo, (int)hi, x);
                L.own = hi > lo ? ((hi == 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u)) : 0u;
            }
        }
        L.endbit = (te >= cs && te < cs + 32) ? 1u << (int)(te - cs) : 0u;
        int64_t first_row = 0;
        int nrows = 0;
        if (EMIT) L.rows = akn3_lane_rows2(B.off, B.n_rows, r_w0, ws0, lane, first_row, nrows);
        else L.rows = akn3_lane_rows(B.off, B.n_rows, r_w0, ws0, lane);
        akwt_phase1(x, L);
        uint32_t dnn = __shfl_down_sync(0xFFFFFFFFu, L.dn, 1);
        if (lane == 31) dnn = 0;
        akwt_phase2(L, dnn, A.mode);
        if (L.hl & ~L.DEV) akwt_wide(B.text, cs, te, L);
        akwt_summary(L);
        uint32_t upp = __shfl_up_sync(0xFFFFFFFFu, L.up, 1);
        if (lane == 0) upp = 0;
        const uint32_t open = akwt_phase3(L, upp);
        const bool real = lane >= 1 && lane <= 30;
        const int n_t = real ? __popc(L.T) : 0;
        int inc = n_t | ((real ? nrows : 0) << 16);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += y;
        }
        if (!EMIT) {
            if (lane == 31) A.count[wt] = inc & 0xFFFF;
            continue;
        }
        if (!real) continue;
        const int64_t t_at = A.base[wt] + ((inc & 0xFFFF) - n_t);                       // tokens that start before this lane
        const int64_t e_at = t_at - (int64_t)open;                                      // tokens that end before it
        const int64_t row_before = r_w0 + ((inc >> 16) - nrows) - 1;                    // last row that starts before this lane
        akwt_emit_lane(L, cs, B.off, B.n_rows, first_row, nrows, row_before, t_at, e_at, A.begin, A.end, A.cap, A.splits, A.row_flags, st);
    }
    ak_raise(B.result, st);
}
