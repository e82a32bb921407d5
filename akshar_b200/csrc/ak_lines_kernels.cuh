// Kernels of the file front end (cores and rationale: ak_lines.cuh): file bytes on the device -> rows.
//   ak_lines_summ_kernel     text -> one transducer per 1024-byte tile
//   ak_lines_resolve_kernel  the tiles' entry states / row ranks / last kept positions (one CTA: each thread folds a
//                            contiguous stretch of tiles, thread 0 chains the 1024 stretch summaries)
//   ak_lines_emit_kernel     text again -> begin / end of every row (absolute positions)
//   ak_lines_len_kernel, (ak_scan_counts_kernel), ak_lines_gather_kernel   rows copied next to each other + row offsets,
//                            the form every other kernel takes
// Algorithmic bytes per file byte: 3 read (two passes + the gather) + 1 written, + 24 per row.
#pragma once
#include "ak_lines.cuh"

struct AkLinesArgs {
    const uint8_t* text;
    int64_t n;                         // file bytes
    int64_t base0;                     // -(address of the file & 15): tiles are 16-byte aligned in the address space
    int64_t n_tiles;                   // tiles of AKLN_TILE bytes covering positions base0 .. n (n itself included)
    AkLineFn* tile_fn;                 // [n_tiles] summaries, then (resolve) the state BEFORE each tile: s = entry state,
                                       //           cnt0 = rows begun before it (low 32 bits), cnt1 = (high 32 bits), lastk
    int64_t* begin;
    int64_t* end;
    int64_t cap;
    int64_t* result;                   // result[0] = rows
};

__device__ __forceinline__ AkLineFn akl_shfl_up(const AkLineFn& f, int d) {
    AkLineFn r;
    r.s = __shfl_up_sync(0xFFFFFFFFu, f.s, d);
    r.cnt0 = __shfl_up_sync(0xFFFFFFFFu, f.cnt0, d);
    r.cnt1 = __shfl_up_sync(0xFFFFFFFFu, f.cnt1, d);
    r.lastk = __shfl_up_sync(0xFFFFFFFFu, f.lastk, d);
    return r;
}

template <bool EMIT>
__global__ void __launch_bounds__(256) ak_lines_kernel(const AkLinesArgs A) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    uint32_t st = 0;
    for (int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < A.n_tiles; tile += warps) {
        // the lane's 32 bytes (tiles start at a 16-byte boundary of the address space: base0 <= 0)
        const int64_t cs = A.base0 + tile * AKLN_TILE + (int64_t)lane * AKLN_SPAN;
        AkLnLane L;
        {
            uint32_t x[8];
            int64_t lo = -cs, hi = A.n - cs;
            lo = lo < 0 ? 0 : (lo > 32 ? 32 : lo);
            hi = hi < 0 ? 0 : (hi > 32 ? 32 : hi);
            if (lo == 0 && hi == 32) {
                const uint4 v0 = *reinterpret_cast<const uint4*>(A.text + cs);
                const uint4 v1 = *reinterpret_cast<const uint4*>(A.text + cs + 16);
                x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w;
                x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
                L.own = 0xFFFFFFFFu;
            } else {
                akn3_load_edge(A.text, cs, (int)lo, (int)hi, x);
                L.own = hi > lo ? ((hi == 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u)) : 0u;
            }
            L.endbit = (A.n >= cs && A.n < cs + 32) ? 1u << (int)(A.n - cs) : 0u;
            akln_phase1(x, L);
            if (L.WIDE) akln_wide(A.text, cs, A.n, L);
        }
        const AkLineFn mine = akln_summary(L, A.text, cs, A.n);
        // inclusive scan of the lanes' transducers
        AkLineFn inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const AkLineFn up = akl_shfl_up(inc, d);
            if (lane >= d) inc = akl_compose(up, inc);
        }
        if (!EMIT) {
            if (lane == 31) A.tile_fn[tile] = inc;
            continue;
        }
        // what holds before this lane: the tile's entry (from the resolve kernel), then the lanes before me
        const AkLineFn in = A.tile_fn[tile];
        AkLineFn before = akl_shfl_up(inc, 1);
        if (lane == 0) before = akl_identity();
        const int64_t tile_rank = (int64_t)(uint32_t)in.cnt0 | ((int64_t)in.cnt1 << 32);
        const uint32_t state = (before.s >> in.s) & 1u;
        const int64_t rank = tile_rank + (in.s ? before.cnt1 : before.cnt0);
        const int64_t lastk = before.lastk >= 0 ? before.lastk : in.lastk;
        akln_emit(L, A.text, cs, A.n, state, lastk, rank, A.begin, A.end, A.cap, st);
    }
    ak_raise(A.result, st);
}

__global__ void __launch_bounds__(1024) ak_lines_resolve_kernel(const AkLinesArgs A) {
    __shared__ AkLineFn chunk[1024];
    const int t = threadIdx.x;
    const int64_t per = (A.n_tiles + 1023) / 1024;
    const int64_t lo = (int64_t)t * per, hi = lo + per < A.n_tiles ? lo + per : A.n_tiles;
    AkLineFn f = akl_identity();
    for (int64_t i = lo; i < hi; ++i) f = akl_compose(f, A.tile_fn[i]);
    chunk[t] = f;
    __syncthreads();
    if (t == 0) {
        // entry of every stretch: state, rows begun, last kept position -- chained from the start of the file (state 0)
        uint32_t state = 0u;
        int64_t rank = 0, lastk = -1;
        for (int c = 0; c < 1024; ++c) {
            const AkLineFn g = chunk[c];
            AkLineFn e;
            e.s = state;
            e.cnt0 = (int32_t)(uint32_t)(rank & 0xFFFFFFFFll);
            e.cnt1 = (int32_t)(rank >> 32);
            e.lastk = lastk;
            chunk[c] = e;
            rank += state ? g.cnt1 : g.cnt0;
            if (g.lastk >= 0) lastk = g.lastk;
            state = (g.s >> state) & 1u;
        }
        A.result[0] = rank;
    }
    __syncthreads();
    uint32_t state = chunk[t].s;
    int64_t rank = (int64_t)(uint32_t)chunk[t].cnt0 | ((int64_t)chunk[t].cnt1 << 32);
    int64_t lastk = chunk[t].lastk;
    for (int64_t i = lo; i < hi; ++i) {
        const AkLineFn g = A.tile_fn[i];
        AkLineFn e;
        e.s = state;
        e.cnt0 = (int32_t)(uint32_t)(rank & 0xFFFFFFFFll);
        e.cnt1 = (int32_t)(rank >> 32);
        e.lastk = lastk;
        A.tile_fn[i] = e;
        rank += state ? g.cnt1 : g.cnt0;
        if (g.lastk >= 0) lastk = g.lastk;
        state = (g.s >> state) & 1u;
    }
}

// row lengths (for the scan that gives the row offsets of the packed text)
__global__ void ak_lines_len_kernel(const int64_t* begin, const int64_t* end, int64_t cap, int32_t* len, int64_t* result) {
    uint32_t st = 0;
    int64_t n_rows = result[0];
    if (n_rows > cap) { n_rows = 0; st |= AK_ST_OVERFLOW; }                   // the caller calls again with room for result[0] rows
    if (blockIdx.x == 0 && threadIdx.x == 0) result[3] = n_rows;             // what the scan and the gather work on
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = end[r] - begin[r];
        if (l < 0 || l > 0x7FFFFFFFll) { st |= AK_ST_OVERFLOW; len[r] = 0; }
        else len[r] = (int32_t)l;
    }
    ak_raise(result, st);
}

// one warp per row: the row's bytes to their place in the packed text
__global__ void __launch_bounds__(256) ak_lines_gather_kernel(const uint8_t* text, const int64_t* begin, const int32_t* len, int64_t* off,
                                                              uint8_t* out, int64_t out_cap, int64_t* result) {
    const int lane = threadIdx.x & 31;
    const int64_t n_rows = result[3];
    if (blockIdx.x == 0 && threadIdx.x == 0) off[n_rows] = result[1];
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    uint32_t st = 0;
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps) {
        const int64_t b = begin[r], o = off[r];
        const int64_t n = len[r];
        if (o + n > out_cap) { st |= AK_ST_OVERFLOW; continue; }
        for (int64_t i = lane; i < n; i += 32) out[o + i] = text[b + i];
    }
    ak_raise(result, st);
}

// rows written one after the other, each followed by `sep` (the output file of preprocess_corpus: line + '\n')
__global__ void __launch_bounds__(256) ak_join_rows_kernel(const uint8_t* text, const int64_t* off, int64_t n_rows, uint8_t sep, uint8_t* out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps) {
        const int64_t b = off[r], n = off[r + 1] - b, o = b - off[0] + r;
        for (int64_t i = lane; i < n; i += 32) out[o + i] = text[b + i];
        if (lane == 0) out[o + n] = sep;
    }
}
