// libakshar_b200.so -- CUDA kernels (sm_100a) + the C ABI declared in include/akshar_b200.h.
//
// Every batch kernel is persistent (gridDim = SMs x resident CTAs); a CTA draws tile numbers from an atomic
// ticket, each thread walks one byte span of the tile (ak_text_core.cuh / ak_subword.cuh), output positions
// come from a block scan plus the single-pass ordered tile prefix in ak_scan.cuh, so the text is read from HBM
// once and every output is written once.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/akshar_b200.h"
#include "ak_models.h"
#include "ak_scan.cuh"
#include "ak_subword.cuh"
#include "ak_fast.cuh"
#include "ak_norm3.cuh"
#include "ak_seg3.cuh"
#include "ak_bpe3.cuh"
#include "ak_bpe_fast.cuh"
#include "ak_seg_fast.cuh"
#include "unicode_tables.inc"

#define AK_BLOCK 256
#define AK_SPAN 32
#define AK_TILE (AK_BLOCK * AK_SPAN)
#define AK_LOOKBACK_LIMIT 4096          // bytes a span may walk backwards in AKSHAR_MODE_TILES
#define AK_BPE_STAGE 40                 // ids per thread staged in shared memory (>= AK_SPAN + a few <s> / </s>)
#define AKB_STAGE 24                    // same, fast kernel (16-byte chunks)
#define AKB_EVCAP 512                   // events (row starts + word starts) per warp tile kept in shared memory
#define AKW_GROUP 256                   // warp tiles per scan group (one CTA of the sums / copy kernels)
#define AKS_STAGE 18                    // cluster / run ends per lane staged in shared memory (fast segment kernel)
#define AK_ROWS_BLOCK 128               // rows per tile for the row-per-thread kernels
#define AKF_WARPS (AK_BLOCK / 32)
#define AKF_TILE (AKF_WARPS * AKF_WARP_BYTES)     // 3840 text bytes per CTA tile in the fast kernels
#define AKF_STAGE (AKF_TILE + 1280)               // shared-memory output stage (normalize can expand a little)

static_assert((int)AK_ST_OVERFLOW == (int)AKSHAR_ST_OVERFLOW && (int)AK_ST_NFC_SEGMENT == (int)AKSHAR_ST_NFC_SEGMENT &&
              (int)AK_ST_PATHOLOGICAL == (int)AKSHAR_ST_PATHOLOGICAL && (int)AK_ST_ALPHABET == (int)AKSHAR_ST_ALPHABET &&
              (int)AK_ST_SPIN == (int)AKSHAR_ST_SPIN && (int)AK_ST_WORD == (int)AKSHAR_ST_WORD, "status bits out of sync");
static_assert(AK_NORM_ROMAN == AKSHAR_NORM_ROMAN && AK_NORM_CLEAN == AKSHAR_NORM_CLEAN && AK_NORM_FILTER == AKSHAR_NORM_FILTER &&
              AK_NORM_COLLAPSE == AKSHAR_NORM_COLLAPSE && AK_NORM_NO_NFC == AKSHAR_NORM_NO_NFC, "flags out of sync");
static_assert(AK_SEG_CLUSTERS == AKSHAR_SEG_CLUSTERS && AK_SEG_MATRAS == AKSHAR_SEG_MATRAS &&
              AK_SEG_RUNS == AKSHAR_SEG_RUNS, "flags out of sync");

// ------------------------------------------------------------------------------------------------
// common kernel plumbing
// ------------------------------------------------------------------------------------------------
struct AkBatch {
    const uint8_t* text;
    const int64_t* off;
    int64_t n_rows, text_begin, text_end;
    int mode;
    int n_tiles;
    int* ticket;
    unsigned long long* state0;
    unsigned long long* state1;
    int64_t* result;           // [4]; status bits are OR-ed into result[2]
    int64_t* totals;           // [2]; normally == result
    const unsigned int* run_if;   // non-null: the kernel is a no-op unless *run_if != 0
    const int64_t* dyn_end;       // non-null: text_end = text_begin + *dyn_end (length produced by an earlier kernel)
};

// start-of-kernel resolution of the device-side conditionals; false = nothing to do
__device__ __forceinline__ bool ak_batch_begin(AkBatch& B) {
    if (B.run_if && *B.run_if == 0) return false;
    if (B.dyn_end) {
        // second stage of a pipeline: the first stage's output is unusable once it gave up or overflowed (the host
        // re-runs the whole call), so do not walk over it
        if (B.result[2] & (AK_ST_OVERFLOW | AK_ST_PATHOLOGICAL | AK_ST_NFC_SEGMENT | AK_ST_SPIN)) return false;
        B.text_end = B.text_begin + *B.dyn_end;
        if (B.mode == AKSHAR_MODE_TILES) B.n_tiles = (int)((B.text_end - B.text_begin + AK_TILE) / AK_TILE);
    }
    return true;
}

struct AkSpan {
    int64_t s, e, r_lo, r_hi, limit;
};

__device__ __forceinline__ void ak_raise(int64_t* result, uint32_t bits) {
    if (bits) atomicOr((unsigned long long*)&result[2], (unsigned long long)bits);
}

// span of this thread inside `tile`; sh[0..1] is CTA scratch for the tile's row window
__device__ __forceinline__ AkSpan ak_span_of(const AkBatch& B, int tile, int64_t* sh) {
    AkSpan sp;
    if (B.mode == AKSHAR_MODE_TILES) {
        const int64_t t0 = B.text_begin + (int64_t)tile * AK_TILE;
        int64_t t1 = t0 + AK_TILE;
        if (t1 > B.text_end + 1) t1 = B.text_end + 1;
        if (threadIdx.x == 0) {
            int64_t lo = ak_row_lower_bound(B.off, 0, B.n_rows, t0);
            sh[0] = lo > 0 ? lo - 1 : 0;
            sh[1] = ak_row_lower_bound(B.off, lo, B.n_rows, t1);
        }
        __syncthreads();
        sp.r_lo = sh[0];
        sp.r_hi = sh[1];
        sp.s = t0 + (int64_t)threadIdx.x * AK_SPAN;
        sp.e = sp.s + AK_SPAN;
        if (sp.e > t1) sp.e = t1;
        if (sp.s > sp.e) sp.s = sp.e;
        sp.limit = AK_LOOKBACK_LIMIT;
    } else {
        const int64_t r = (int64_t)tile * AK_BLOCK + threadIdx.x;
        sp.r_lo = 0;
        sp.r_hi = B.n_rows;
        sp.limit = 0;
        if (r < B.n_rows) {
            sp.s = B.off[r];
            sp.e = (r == B.n_rows - 1) ? B.text_end + 1 : B.off[r + 1];
        } else {
            sp.s = sp.e = 0;
        }
    }
    return sp;
}

__device__ __forceinline__ int ak_next_tile(int* ticket, int* sh) {
    __syncthreads();
    if (threadIdx.x == 0) *sh = atomicAdd(ticket, 1);
    __syncthreads();
    return *sh;
}

// ------------------------------------------------------------------------------------------------
// K1 normalize_text  (reference normalize.py:117-148)
// ------------------------------------------------------------------------------------------------
struct AkNormArgs {
    AkBatch B;
    AkTables T;
    uint32_t flags;
    uint8_t* out;
    int64_t out_cap;
    int64_t* out_off;
};

__global__ void __launch_bounds__(AK_BLOCK) ak_normalize_kernel(const AkNormArgs A) {
    __shared__ int ws[33];
    __shared__ int s_tile;
    __shared__ int64_t s_win[2];
    __shared__ long long s_base;
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    for (;;) {
        const int tile = ak_next_tile(B.ticket, &s_tile);
        if (tile >= B.n_tiles) break;
        const AkSpan sp = ak_span_of(B, tile, s_win);
        uint32_t st = 0;
        int cnt = 0;
        if (sp.s < sp.e)
            cnt = (int)ak_norm_span(A.T, B.text, B.off, B.n_rows, sp.r_lo, sp.r_hi, sp.s, sp.e, A.flags, sp.limit, nullptr,
                                    nullptr, 0, st);
        int total;
        const int pre = ak_block_exscan<AK_BLOCK>(cnt, ws, total);
        if (threadIdx.x < 32) {
            long long b = ak_tile_prefix(B.state0, tile, total, (unsigned int*)&B.result[2], AK_ST_SPIN);
            if (threadIdx.x == 0) {
                s_base = b;
                if (tile == B.n_tiles - 1) B.totals[0] = b + total;
            }
        }
        __syncthreads();
        const int64_t obase = s_base + pre;
        if (sp.s < sp.e) {
            uint8_t* o = nullptr;
            if (obase + cnt <= A.out_cap) o = A.out + obase;
            else if (cnt > 0) st |= AK_ST_OVERFLOW;
            uint32_t st2 = 0;
            ak_norm_span(A.T, B.text, B.off, B.n_rows, sp.r_lo, sp.r_hi, sp.s, sp.e, A.flags, sp.limit, o, A.out_off, obase, st2);
        }
        ak_raise(B.result, st);
    }
}

// ------------------------------------------------------------------------------------------------
// K1 fast: normalize_text with the default flags (NFC + Roman lowercase + allow-list + elongation collapse).
// 16 bytes per thread in registers, emit-mask fast lane (ak_fast.cuh), exact walker as the per-thread slow lane,
// shared-memory output stage flushed with 16-byte stores, row offsets from per-chunk prefix + emit mask.
// ------------------------------------------------------------------------------------------------
struct AkFastNormArgs {
    AkBatch B;
    AkTables T;
    uint8_t* out;
    int64_t out_cap;
    int64_t* out_off;
    const int64_t* tile_row;     // [n_tiles + 1]: first row r in [0, n_rows] with off[r] >= start of tile k (n_rows + 1 if none)
    int64_t base0;               // 16-byte aligned (as an address) start of tile 0, <= text_begin
    uint32_t flags;              // AK_NORM_ROMAN | AK_NORM_CLEAN, or AK_NORM_ROMAN alone (clean_hinglish=False; bit-stream kernel only)
};

// a chunk that straddles the start / end of the text: byte by byte, guarded.  Cold, kept out of line.
__device__ __noinline__ void akf_load_edge(const uint8_t* text, int64_t cs, int lo, int hi, uint32_t* w) {
    w[0] = w[1] = w[2] = w[3] = 0;
#pragma unroll 1
    for (int i = lo; i < hi; ++i) w[i >> 2] |= (uint32_t)text[cs + i] << ((i & 3) * 8);
}

template <class CH>
__device__ __forceinline__ void akf_load_chunk(const uint8_t* text, int64_t cs, int64_t tb, int64_t te, CH& c) {
    int64_t lo = tb - cs, hi = te - cs;
    lo = lo < 0 ? 0 : (lo > 16 ? 16 : lo);
    hi = hi < 0 ? 0 : (hi > 16 ? 16 : hi);
    c.own = hi > lo ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
    if (c.own == 0xFFFFu) {
        const uint4 v = *reinterpret_cast<const uint4*>(text + cs);
        c.w[0] = v.x; c.w[1] = v.y; c.w[2] = v.z; c.w[3] = v.w;
    } else {
        uint32_t w[4];
        akf_load_edge(text, cs, (int)lo, (int)hi, w);
        c.w[0] = w[0]; c.w[1] = w[1]; c.w[2] = w[2]; c.w[3] = w[3];
    }
}

// ------------------------------------------------------------------------------------------------
// Warp tiles.  The fast BPE and segment kernels are warp-autonomous: a warp owns 480 text bytes (30 real lanes +
// 2 halo lanes), finds the rows that start in them with shuffles, encodes, and appends its output to its CTA's
// private slice of a temporary stream (cursor in shared memory) -- no CTA barrier and no global atomic on the hot
// path, so a slow lane (cache miss, long word, slow-lane walker) only delays its own warp.  A scan over the
// per-warp-tile totals then gives the final positions and a copy kernel moves the blocks.
// ------------------------------------------------------------------------------------------------

// wrow[k] = first row r in [0, n_rows] with off[r] >= base0 + k * 480 (n_rows + 1 if none); one thread per entry
__global__ void ak_warp_rows_kernel(AkBatch B, int64_t base0, int n_entries, int64_t* wrow) {
    if (!ak_batch_begin(B)) return;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_entries) return;
    const int64_t pos = base0 + (int64_t)k * AKF_WARP_BYTES;
    int64_t r = ak_row_lower_bound(B.off, 0, B.n_rows, pos);
    if (B.off[r] < pos) r = B.n_rows + 1;
    wrow[k] = r;
}

__device__ __forceinline__ int akw_n_tiles(const AkBatch& B, int64_t base0) {
    return (int)((B.text_end - base0 + AKF_WARP_BYTES) / AKF_WARP_BYTES);      // covers position text_end itself
}

// each lane's 16-bit row-start mask for its chunk [cs, cs + 16), from the sorted row offsets (no shared memory)
__device__ __forceinline__ uint32_t akw_lane_rows(const int64_t* off, int64_t n_rows, int64_t r_w0, int64_t ws, int lane) {
    uint32_t rows = 0;
    const int64_t lo = ws - 16, hi = ws + AKF_WARP_BYTES + 16;      // positions of lanes 0 .. 31
    for (int64_t r = r_w0;; r += 32) {                              // rows at or after ws
        const int64_t mr = r + lane;
        const int64_t p = mr <= n_rows ? off[mr] : hi;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, p < hi);
        const int cnt = __popc(m);                                  // sorted: the in-range rows are a prefix
        for (int j = 0; j < cnt; ++j) {
            const int64_t pj = __shfl_sync(0xFFFFFFFFu, p, j);
            const int rel = (int)(pj - lo);
            if ((rel >> 4) == lane) rows |= 1u << (rel & 15);
        }
        if (cnt < 32) break;
    }
    for (int64_t r = r_w0 - 1;; r -= 32) {                          // rows inside the left halo chunk
        const int64_t mr = r - lane;
        const int64_t p = mr >= 0 ? off[mr] : lo - 1;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, p >= lo);
        const int cnt = __popc(m);
        if (cnt && lane == 0) {
            // all of them fall into lane 0's chunk
        }
        for (int j = 0; j < cnt; ++j) {
            const int64_t pj = __shfl_sync(0xFFFFFFFFu, p, j);
            const int rel = (int)(pj - lo);
            if ((rel >> 4) == lane) rows |= 1u << (rel & 15);
        }
        if (cnt < 32) break;
    }
    return rows;
}

// sums of AKW_GROUP consecutive warp-tile totals
__global__ void __launch_bounds__(AKW_GROUP) ak_wt_sums_kernel(AkBatch B, int64_t base0, const int32_t* wt_total, int32_t* sums) {
    __shared__ int ws[33];
    if (!ak_batch_begin(B)) return;
    const int n_wt = akw_n_tiles(B, base0);
    const int n_groups = (n_wt + AKW_GROUP - 1) / AKW_GROUP;
    for (int gidx = blockIdx.x; gidx < n_groups; gidx += gridDim.x) {
        const int t = gidx * AKW_GROUP + threadIdx.x;
        int total;
        ak_block_exscan<AKW_GROUP>(t < n_wt ? wt_total[t] : 0, ws, total);
        if (threadIdx.x == 0) sums[gidx] = total;
    }
}

// Slow chunks are not processed where they are found: a lane that cannot take the fast lane appends its chunk to a
// work list, and two small kernels run the exact walker over that list with one thread per entry.  A 16-byte walk
// costs tens of microseconds of dependent instructions; inside the tile kernels it would stall its whole CTA (and,
// through an ordered tile prefix, every later tile), on the list thousands of them overlap.
#define AK_SLOW_BYTES 72
struct AkSlowEntry {
    int64_t pos;         // span start (absolute byte index)
    int64_t out_base;    // filled by the write kernel: where this span's output starts
    int32_t cnt;         // filled by the slow kernel's first pass
    int32_t tile;
    int32_t span;        // 16, or 32: both chunks of a bit-parallel lane in one walk (the second chunk's info word is
    int32_t pad_;        // 0xC0000000 | index: "continued", no bytes of its own)
    uint8_t bytes[AK_SLOW_BYTES];      // the span's output when it fits (else the second pass walks again)
};

struct AkNfWork {
    uint32_t* info;            // [n_tiles * AK_BLOCK] per lane: emit mask, or 0x80000000 | work-list index
    int32_t* tile_total;       // [n_tiles] output bytes of the tile
    int64_t* tile_base;        // [n_tiles + 1] exclusive prefix
    AkSlowEntry* slow;
    unsigned int* n_slow;
    unsigned int slow_cap;
};

// chunk bytes + the 4 bytes that follow (from the next lane; the right halo reads them itself)
template <class CH>
__device__ __forceinline__ void akf_load_lane(const AkBatch& B, int64_t cs, CH& c) {
    akf_load_chunk(B.text, cs, B.text_begin, B.text_end, c);
    uint32_t nx = __shfl_down_sync(0xFFFFFFFFu, c.w[0], 1);
    if ((threadIdx.x & 31) == 31) {
        nx = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t q = cs + 16 + i;
            if (q >= B.text_begin && q < B.text_end) nx |= (uint32_t)B.text[q] << (i * 8);
        }
    }
    c.w[4] = nx;
}

// ---- K1a: classify every chunk: emit mask for the fast lane, work-list entry otherwise; per-tile fast byte counts
#ifndef AKN_MINB
#define AKN_MINB 4
#endif
__global__ void __launch_bounds__(AK_BLOCK, AKN_MINB) ak_nf_classify_kernel(const AkFastNormArgs A, const AkNfWork W) {
    __shared__ uint32_t lut[384];
    __shared__ int s_red[AKF_WARPS];
    const AkBatch& B = A.B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 384; i += AK_BLOCK)
        lut[i] = i < 128 ? A.T.leaves[((uint32_t)A.T.page_index[0] << 8) | i] : A.T.leaves[((uint32_t)A.T.page_index[9] << 8) | (i - 128)];
    __syncthreads();
    for (int tile = blockIdx.x; tile < B.n_tiles; tile += gridDim.x) {
        const int64_t tile_start = A.base0 + (int64_t)tile * AKF_TILE;
        const int64_t ws = tile_start + (int64_t)warp * AKF_WARP_BYTES;
        AkChunk c;
        const int64_t cs = ws + (int64_t)(lane - 1) * 16;
        akf_load_lane(B, cs, c);
        // row starts of the warp's 32 chunks straight from the sorted offsets (tile_row has one entry per warp tile)
        c.rows = akw_lane_rows(B.off, B.n_rows, A.tile_row[(size_t)tile * AKF_WARPS + warp], ws, lane);
        if (c.own == 0 && c.rows == 0) {
            // entirely outside the text: acts as a row boundary for its neighbours
            c.kept = c.lead = 0;
            c.flags = AKF_BOUNDARY | AKF_ROWSTART;
            c.first_w = c.last_w = c.F = c.L1 = c.L2 = AKF_NONE;
        } else {
            akf_phase_a(A.T, lut, c);
        }
        // the first owned code point against the last one of the previous chunk
        {
            uint32_t pl = __shfl_up_sync(0xFFFFFFFFu, c.last_w, 1);
            if (lane == 0) {
                // left halo: decode the code point that ends right before the chunk (same row only)
                pl = AKF_NONE;
                if (c.first_w != AKF_NONE && cs > B.text_begin) {
                    int64_t q = cs - 1;
                    int k = 0;
                    while (q > B.text_begin && k < 3 && (B.text[q] & 0xC0u) == 0x80u) { --q; ++k; }
                    int len;
                    pl = akf_props(A.T, lut, ak_decode(B.text, q, B.text_end, len));
                }
            }
            akf_resolve_first(c, pl);
        }
        AkNeighbor pv, nx;
        pv.flags = __shfl_up_sync(0xFFFFFFFFu, c.flags, 1);
        pv.F = AKF_NONE;
        pv.L1 = __shfl_up_sync(0xFFFFFFFFu, c.L1, 1);
        pv.L2 = __shfl_up_sync(0xFFFFFFFFu, c.L2, 1);
        nx.flags = __shfl_down_sync(0xFFFFFFFFu, c.flags, 1);
        nx.F = __shfl_down_sync(0xFFFFFFFFu, c.F, 1);
        nx.L1 = nx.L2 = AKF_NONE;
        const bool real = lane >= 1 && lane <= AKF_REAL;
        const int64_t ss = cs < B.text_begin ? B.text_begin : cs;
        const int64_t se = cs + 16 > B.text_end + 1 ? B.text_end + 1 : cs + 16;
        uint32_t info = 0;
        int cnt = 0;
        if (real && ss < se) {
            uint32_t emit = 0;
            const bool slow = akf_is_slow(c, pv, nx) || !akf_collapse(c, pv, nx, emit);
            if (slow) {
                const unsigned int idx = atomicAdd(W.n_slow, 1u);
                if (idx < W.slow_cap) {
                    AkSlowEntry e;
                    e.pos = cs;
                    e.out_base = 0;
                    e.cnt = 0;
                    e.tile = tile;
                    e.span = 16;
                    e.pad_ = 0;
                    W.slow[idx] = e;
                } else {
                    ak_raise(B.result, AK_ST_PATHOLOGICAL);     // too many slow chunks: the host re-runs row by row
                }
                info = 0x80000000u | idx;
            } else {
                info = emit;
                cnt = __popc(emit);
            }
        }
        W.info[(size_t)tile * AK_BLOCK + tid] = info;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        if (lane == 0) s_red[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
#pragma unroll
            for (int w = 0; w < AKF_WARPS; ++w) t += s_red[w];
            W.tile_total[tile] = t;
        }
        __syncthreads();       // s_red is reused by the next tile
    }
}

// ---- K1a v3: the same classification as parallel bit streams (ak_norm3.cuh): 32 bytes per lane, 30 real lanes per
// warp (960 bytes = two 480-byte warp tiles of the v2 geometry), 4 warps per 3840-byte tile.  Produces exactly what
// ak_nf_classify_kernel produces -- one 19-bit emit mask per 16-byte chunk, work-list entries for the slow chunks,
// the tile's fast byte count -- so the scan / write / slow kernels are shared.
#define AKN3_THREADS 128
#define AKN3_WARP_BYTES 960

// row-start mask of the lane's 32 bytes [ws + 32 (lane - 1), +32) from the sorted offsets; r_w0 = first row at or after ws
__device__ __forceinline__ uint32_t akn3_lane_rows(const int64_t* off, int64_t n_rows, int64_t r_w0, int64_t ws, int lane) {
    uint32_t rows = 0;
    const int64_t lo = ws - 32, hi = ws + AKN3_WARP_BYTES + 32;
    for (int64_t r = r_w0;; r += 32) {
        const int64_t mr = r + lane;
        const int64_t p = mr <= n_rows ? off[mr] : hi;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, p < hi);
        const int cnt = __popc(m);
        for (int j = 0; j < cnt; ++j) {
            const int rel = (int)(__shfl_sync(0xFFFFFFFFu, p, j) - lo);
            if ((rel >> 5) == lane) rows |= 1u << (rel & 31);
        }
        if (cnt < 32) break;
    }
    for (int64_t r = r_w0 - 1;; r -= 32) {
        const int64_t mr = r - lane;
        const int64_t p = mr >= 0 ? off[mr] : lo - 1;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, p >= lo);
        const int cnt = __popc(m);
        for (int j = 0; j < cnt; ++j) {
            const int rel = (int)(__shfl_sync(0xFFFFFFFFu, p, j) - lo);
            if ((rel >> 5) == lane) rows |= 1u << (rel & 31);
        }
        if (cnt < 32) break;
    }
    return rows;
}

__device__ __noinline__ void akn3_load_edge(const uint8_t* text, int64_t cs, int lo, int hi, uint32_t* x) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0;
#pragma unroll 1
    for (int i = lo; i < hi; ++i) x[i >> 2] |= (uint32_t)text[cs + i] << ((i & 3) * 8);
}

#ifndef AKN3_MINB
#define AKN3_MINB 8
#endif
__global__ void __launch_bounds__(AKN3_THREADS, AKN3_MINB) ak_nf3_classify_kernel(const AkFastNormArgs A, const AkNfWork W) {
    __shared__ int s_red[AKN3_THREADS / 32];
    const AkBatch& B = A.B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tb = B.text_begin, te = B.text_end;
    const bool raw = !(A.flags & AK_NORM_CLEAN);
    for (int tile = blockIdx.x; tile < B.n_tiles; tile += gridDim.x) {
        const int64_t tile_start = A.base0 + (int64_t)tile * AKF_TILE;
        const int64_t ws = tile_start + (int64_t)warp * AKN3_WARP_BYTES;
        const int64_t cs = ws + (int64_t)(lane - 1) * 32;
        AkN3Lane L;
        {
            uint32_t x[8];
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
            L.rows = akn3_lane_rows(B.off, B.n_rows, A.tile_row[(size_t)tile * AKF_WARPS + 2 * warp], ws, lane);
            akn3_phase1(x, L);
        }
        uint32_t up1p = __shfl_up_sync(0xFFFFFFFFu, L.up1, 1);
        uint32_t dn1n = __shfl_down_sync(0xFFFFFFFFu, L.dn1, 1);
        if (lane == 0) up1p = 0;
        if (lane == 31) dn1n = 0;
        akn3_phase2(L, up1p, dn1n, raw);
        uint32_t up2p = __shfl_up_sync(0xFFFFFFFFu, L.up2, 1);
        uint32_t dn2n = __shfl_down_sync(0xFFFFFFFFu, L.dn2, 1);
        if (lane == 0) up2p = AKN3_HALO_UP2;
        if (lane == 31) dn2n = 0;
        akn3_phase3(A.T, B.text, cs, te, L, up2p, dn2n, raw);
        {
            uint32_t rest = 0;
            if (L.ge) rest = akn3_gaps_local(B.text, cs, te, L);
            if (lane == 0) {                                       // the halo lane cannot look further left
                akn3_gaps_remote(B.text, cs, te, L, rest, 0u);
                rest = 0;
            }
            if (__any_sync(0xFFFFFFFFu, rest != 0u)) {
                const uint32_t lk = akn3_last_kept(B.text, cs, te, L);
                const uint32_t plk = __shfl_up_sync(0xFFFFFFFFu, lk, 1);
                akn3_gaps_remote(B.text, cs, te, L, rest, plk);
            }
        }
        akn3_phase3b(L);
        const uint32_t up3p = __shfl_up_sync(0xFFFFFFFFu, L.up3, 1);
        const uint32_t dn3n = __shfl_down_sync(0xFFFFFFFFu, L.dn3, 1);
        uint32_t info[2] = {0u, 0u};
        const bool fast = akn3_phase4(L, up3p, dn1n, dn3n, info[0], info[1]);
        int cnt = 0;
        if (lane >= 1 && lane <= 30) {
            bool act[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t hs = cs + 16 * h;
                const int64_t ss = hs < tb ? tb : hs;
                const int64_t se = hs + 16 > te + 1 ? te + 1 : hs + 16;
                act[h] = ss < se;
            }
            uint32_t v[2] = {0u, 0u};
            if (fast) {
                if (act[0]) { v[0] = info[0]; cnt += __popc(v[0]); }
                if (act[1]) { v[1] = info[1]; cnt += __popc(v[1]); }
            } else if (act[0] || act[1]) {
                // one work-list entry for the lane: both chunks in one walk
                const unsigned int idx = atomicAdd(W.n_slow, 1u);
                if (idx < W.slow_cap) {
                    AkSlowEntry e;
                    e.pos = act[0] ? cs : cs + 16;
                    e.out_base = 0;
                    e.cnt = 0;
                    e.tile = tile;
                    e.span = (act[0] && act[1]) ? 32 : 16;
                    e.pad_ = 0;
                    W.slow[idx] = e;
                } else {
                    ak_raise(B.result, AK_ST_PATHOLOGICAL);
                }
                if (act[0]) { v[0] = 0x80000000u | idx; if (act[1]) v[1] = 0xC0000000u | idx; }
                else v[1] = 0x80000000u | idx;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = (int)((cs + 16 * h - tile_start) >> 4);                 // 16-byte chunk of the tile, 0 .. 239
                W.info[(size_t)tile * AK_BLOCK + (k / AKF_REAL) * 32 + 1 + (k % AKF_REAL)] = v[h];
            }
        }
        if (tid < 2 * AKF_WARPS) W.info[(size_t)tile * AK_BLOCK + (tid >> 1) * 32 + (tid & 1) * 31] = 0;   // the v2 halo slots
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        if (lane == 0) s_red[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
#pragma unroll
            for (int w = 0; w < AKN3_THREADS / 32; ++w) t += s_red[w];
            W.tile_total[tile] = t;
        }
        __syncthreads();
    }
}

// ---- K1b / K1e: the walker over the work list (count pass, then write pass)
struct AkNfSlowArgs {
    AkBatch B;
    AkTables T;
    AkNfWork W;
    const int64_t* tile_row;
    uint8_t* out;
    int64_t out_cap;
    int64_t* out_off;
    int write;
    uint32_t flags;
};

#ifndef AKN_SLOW_MINB
#define AKN_SLOW_MINB 12       // measured 4 / 6 / 8 / 12: 8.18 / 8.02 / 8.01 / 7.89 ms per 512 MiB of the BPE workload
#endif
__global__ void __launch_bounds__(128, AKN_SLOW_MINB) ak_nf_slow_kernel(const AkNfSlowArgs A) {
    const AkBatch& B = A.B;
    unsigned int n = *A.W.n_slow;
    if (n > A.W.slow_cap) n = A.W.slow_cap;
    const uint32_t NFLAGS = A.flags;
    for (unsigned int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        AkSlowEntry e = A.W.slow[j];
        const int64_t ss = e.pos < B.text_begin ? B.text_begin : e.pos;
        const int64_t se = e.pos + e.span > B.text_end + 1 ? B.text_end + 1 : e.pos + e.span;
        const int64_t r0 = A.tile_row[(size_t)e.tile * AKF_WARPS], r1 = A.tile_row[(size_t)(e.tile + 1) * AKF_WARPS];
        const int64_t rlo = r0 > 0 ? r0 - 1 : 0, rhi = r1 > B.n_rows ? B.n_rows : r1;
        uint32_t st = 0;
        if (!A.write) {
            // one walk: bytes into the entry, row offsets relative to the chunk's output (the write kernel rebases them)
            const int cnt = (int)ak_norm_span(A.T, B.text, B.off, B.n_rows, rlo, rhi, ss, se, NFLAGS, AK_LOOKBACK_LIMIT,
                                              A.W.slow[j].bytes, A.out_off, 0, st, AK_SLOW_BYTES);
            A.W.slow[j].cnt = cnt;
            atomicAdd(&A.W.tile_total[e.tile], cnt);
        } else if (e.cnt > AK_SLOW_BYTES) {
            uint8_t* dst = (e.out_base + e.cnt <= A.out_cap) ? A.out + e.out_base : nullptr;
            ak_norm_span(A.T, B.text, B.off, B.n_rows, rlo, rhi, ss, se, NFLAGS, AK_LOOKBACK_LIMIT, dst, A.out_off, e.out_base, st);
        }
        ak_raise(B.result, st);
    }
}

// ---- K1c: exclusive prefix of the tile totals (one CTA; the array has one entry per 3840 bytes of text)
__global__ void __launch_bounds__(1024) ak_nf_scan_kernel(const int32_t* tile_total, int64_t* tile_base, int n_tiles,
                                                          int64_t* total_out, AkBatch B, int64_t base0,
                                                          int64_t bytes_per_entry = AKF_TILE) {
    __shared__ long long ws[33];
    __shared__ long long carry;
    if (!ak_batch_begin(B)) return;
    if (B.dyn_end) {
        if (bytes_per_entry == AKF_TILE) n_tiles = (int)((B.text_end - base0 + AKF_TILE) / AKF_TILE);
        else {
            const int n_wt = (int)((B.text_end - base0 + AKF_WARP_BYTES) / AKF_WARP_BYTES);
            n_tiles = (n_wt + AKW_GROUP - 1) / AKW_GROUP;
        }
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    // 8 consecutive entries per thread (a serial prefix in registers), so one trip of the block scan covers 8192 entries
    for (int b = 0; b < n_tiles; b += 8192) {
        const int i0 = b + tid * 8;
        int v8[8];
        long long v = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v8[k] = (i0 + k < n_tiles) ? tile_total[i0 + k] : 0;
            v += v8[k];
        }
        long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            long long y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += y;
        }
        if (lane == 31) ws[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long x = ws[lane], xi = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                long long y = __shfl_up_sync(0xFFFFFFFFu, xi, d);
                if (lane >= d) xi += y;
            }
            ws[lane] = xi - x;
            if (lane == 31) ws[32] = xi;
        }
        __syncthreads();
        long long ex = carry + ws[warp] + inc - v;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (i0 + k < n_tiles) tile_base[i0 + k] = ex;
            ex += v8[k];
        }
        __syncthreads();
        if (tid == 0) carry += ws[32];
        __syncthreads();
    }
    if (tid == 0) {
        tile_base[n_tiles] = carry;
        *total_out = carry;
    }
}

// The writer's common case: the emitted bytes of a chunk are ONE contiguous stretch of its 20-byte window (nothing dropped
// inside; the first bytes may belong to the previous chunk's last code point, the last code point may reach into the next
// chunk).  A-Z lowered four bytes at a time, the stretch moved with funnel shifts: bytes up to the destination's next word
// boundary one by one, then whole words, then the tail.
__device__ __forceinline__ void akf_write_run(const AkChunk& c, uint32_t emit, uint8_t* dst) {
    uint32_t w[7];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const uint32_t x = c.w[j];
        const uint32_t t7 = x & 0x7F7F7F7Fu;
        const uint32_t up = ((t7 + 0x3F3F3F3Fu) & ~(t7 + 0x25252525u) & ~x) & 0x80808080u;      // 0x41 .. 0x5A
        w[j] = x | (up >> 2);
    }
    w[5] = w[6] = 0;
    const int a = __ffs(emit) - 1, n = __popc(emit);
    int h = (int)((4u - ((uint32_t)(uintptr_t)dst & 3u)) & 3u);
    if (h > n) h = n;
    const int t = a + h;                                  // window byte where the word-aligned part starts (0 .. 6)
    const uint32_t sh = (uint32_t)(t & 3) * 8u;
    uint32_t x[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) x[j] = t >= 4 ? w[j + 1] : w[j];
    uint32_t f[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) f[j] = __funnelshift_r(x[j], x[j + 1], sh);
    {   // head: window bytes a .. a + h - 1
        const uint32_t hv = __funnelshift_r(w[0], w[1], (uint32_t)a * 8u);      // a <= 3
        if (h > 0) dst[0] = (uint8_t)hv;
        if (h > 1) dst[1] = (uint8_t)(hv >> 8);
        if (h > 2) dst[2] = (uint8_t)(hv >> 16);
    }
    const int nw = (n - h) >> 2, r = (n - h) & 3;
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + h);
#pragma unroll
    for (int m = 0; m < 4; ++m) if (m < nw) d32[m] = f[m];
    const uint32_t tv = nw == 0 ? f[0] : nw == 1 ? f[1] : nw == 2 ? f[2] : nw == 3 ? f[3] : f[4];
    uint8_t* dt = dst + h + 4 * nw;
    if (r > 0) dt[0] = (uint8_t)tv;
    if (r > 1) dt[1] = (uint8_t)(tv >> 8);
    if (r > 2) dt[2] = (uint8_t)(tv >> 16);
}

// ---- K1d: write the fast lanes' bytes (staged in shared memory, 16-byte stores) and the row offsets
__global__ void __launch_bounds__(AK_BLOCK) ak_nf_write_kernel(const AkFastNormArgs A, const AkNfWork W) {
    __shared__ __align__(16) uint8_t stage[AKF_STAGE + 32];
    __shared__ uint32_t s_emit[AK_BLOCK];
    __shared__ uint32_t s_pre[AK_BLOCK];
    __shared__ int ws[33];
    const AkBatch& B = A.B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int tile = blockIdx.x; tile < B.n_tiles; tile += gridDim.x) {
        const int64_t tile_start = A.base0 + (int64_t)tile * AKF_TILE;
        const int64_t r0 = A.tile_row[(size_t)tile * AKF_WARPS], r1 = A.tile_row[(size_t)(tile + 1) * AKF_WARPS];
        const int64_t cs = tile_start + (int64_t)warp * AKF_WARP_BYTES + (int64_t)(lane - 1) * 16;
        AkChunk c;
        akf_load_lane(B, cs, c);
        const uint32_t info = W.info[(size_t)tile * AK_BLOCK + tid];
        const bool slow = (info & 0x80000000u) != 0;
        const bool cont = slow && (info & 0x40000000u);          // second chunk of a 32-byte slow span: nothing of its own
        const unsigned int sidx = info & 0x3FFFFFFFu;
        int cnt = 0;
        if (slow) { if (sidx < W.slow_cap && !cont) cnt = W.slow[sidx].cnt; }
        else cnt = __popc(info);
        int total;
        const int pre = ak_block_exscan<AK_BLOCK>(cnt, ws, total);
        const int64_t base = W.tile_base[tile];
        const bool fits = base + total <= A.out_cap;
        const bool staged = fits && total <= AKF_STAGE;
        const int pad = (int)((uintptr_t)(A.out + base) & 15);
        if (!fits && tid == 0 && total > 0) ak_raise(B.result, AK_ST_OVERFLOW);
        s_emit[tid] = info;
        s_pre[tid] = (uint32_t)pre;
        if (slow) {
            if (sidx < W.slow_cap && !cont) {
                W.slow[sidx].out_base = base + pre;
                if (cnt <= AK_SLOW_BYTES && fits) {
                    uint8_t* dst = staged ? stage + pad + pre : A.out + base + pre;
                    const uint8_t* src = W.slow[sidx].bytes;
                    for (int i = 0; i < cnt; ++i) dst[i] = src[i];
                }
            }
        } else if (info && fits) {
            uint8_t* dst = staged ? stage + pad + pre : A.out + base + pre;
            const uint32_t lowbit = info & (0u - info);
            if (staged && ((info + lowbit) & info) == 0u && lowbit <= 8u) akf_write_run(c, info, dst);     // one contiguous stretch
            else akf_write(c, info, dst);
        }
        __syncthreads();
        if (staged) {
            // stage[pad .. pad + total) -> out[base ..): stage and global share their alignment modulo 16.  The holes
            // of slow chunks are copied as garbage here and filled by the slow write kernel afterwards.
            uint8_t* g = A.out + base;
            int head = (16 - pad) & 15;
            if (head > total) head = total;
            if (tid < head) g[tid] = stage[pad + tid];
            const int body = (total - head) >> 4;
            for (int i = tid; i < body; i += AK_BLOCK)
                *reinterpret_cast<uint4*>(g + head + 16 * i) = *reinterpret_cast<const uint4*>(stage + pad + head + 16 * i);
            const int tail0 = head + (body << 4);
            if (tid < total - tail0) g[tail0 + tid] = stage[pad + tail0 + tid];
        }
        // row offsets of the rows that start in a fast chunk of this tile (slow chunks write their own)
        for (int64_t r = r0 + tid; r < r1 && r <= B.n_rows; r += AK_BLOCK) {
            const int rel = (int)(B.off[r] - tile_start);
            const int wq = rel / AKF_WARP_BYTES, within = rel - wq * AKF_WARP_BYTES;
            const int th = wq * 32 + 1 + (within >> 4), i = within & 15;
            const uint32_t e = s_emit[th];
            if (!(e & 0x80000000u)) A.out_off[r] = base + s_pre[th] + __popc(e & ((1u << i) - 1u));
            else {
                // the slow pass left it relative to the span's output; a continued chunk's prefix already includes the span
                int64_t adj = 0;
                if ((e & 0x40000000u) && (e & 0x3FFFFFFFu) < W.slow_cap) adj = W.slow[e & 0x3FFFFFFFu].cnt;
                A.out_off[r] += base + s_pre[th] - adj;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// K2 + K3 grapheme clusters and script runs  (reference segment.py:40-201)
// ------------------------------------------------------------------------------------------------
struct AkSegArgs {
    AkBatch B;
    AkTables T;
    uint32_t flags;
    AkSegOut o;
};

__global__ void __launch_bounds__(AK_BLOCK) ak_segment_kernel(const AkSegArgs A) {
    __shared__ int ws[33];
    __shared__ int s_tile;
    __shared__ int64_t s_win[2];
    __shared__ long long s_base[2];
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    for (;;) {
        const int tile = ak_next_tile(B.ticket, &s_tile);
        if (tile >= B.n_tiles) break;
        const AkSpan sp = ak_span_of(B, tile, s_win);
        uint32_t st = 0;
        int64_t cc = 0, rc = 0;
        AkSegOut o = A.o;
        if (sp.s < sp.e)
            ak_seg_span(A.T, B.text, B.off, B.n_rows, sp.r_lo, sp.r_hi, sp.s, sp.e, A.flags, sp.limit, false, o, cc, rc, st);
        int ctot, rtot;
        const int cpre = ak_block_exscan<AK_BLOCK>((int)cc, ws, ctot);
        const int rpre = ak_block_exscan<AK_BLOCK>((int)rc, ws, rtot);
        if (threadIdx.x < 32) {
            long long cb = ak_tile_prefix(B.state0, tile, ctot, (unsigned int*)&B.result[2], AK_ST_SPIN);
            long long rb = ak_tile_prefix(B.state1, tile, rtot, (unsigned int*)&B.result[2], AK_ST_SPIN);
            if (threadIdx.x == 0) {
                s_base[0] = cb;
                s_base[1] = rb;
                if (tile == B.n_tiles - 1) { B.totals[0] = cb + ctot; B.totals[1] = rb + rtot; }
            }
        }
        __syncthreads();
        if (sp.s < sp.e) {
            o.cbase = s_base[0] + cpre;
            o.rbase = s_base[1] + rpre;
            if (o.cbase + cc > o.ccap || o.rbase + rc > o.rcap) st |= AK_ST_OVERFLOW;
            uint32_t st2 = 0;
            int64_t c2, r2;
            ak_seg_span(A.T, B.text, B.off, B.n_rows, sp.r_lo, sp.r_hi, sp.s, sp.e, A.flags, sp.limit, true, o, c2, r2, st2);
        }
        ak_raise(B.result, st);
    }
}

// ------------------------------------------------------------------------------------------------
// K4a BPE encode  (reference tokenizer.py:193)
// ------------------------------------------------------------------------------------------------
struct AkBpeArgs {
    AkBatch B;
    AkTables T;
    AkBpeDev M;
    AkPool pool;
    int32_t* ids;
    int64_t id_cap;
    int64_t* id_splits;
    unsigned int* changed;          // set when NFC would change the text (results are then recomputed)
};

__global__ void __launch_bounds__(AK_BLOCK) ak_bpe_kernel(const AkBpeArgs A) {
    __shared__ int ws[33];
    __shared__ int s_tile;
    __shared__ int64_t s_win[2];
    __shared__ long long s_base;
    __shared__ int32_t stage[AK_BPE_STAGE * AK_BLOCK];
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    for (;;) {
        const int tile = ak_next_tile(B.ticket, &s_tile);
        if (tile >= B.n_tiles) break;
        const AkSpan sp = ak_span_of(B, tile, s_win);
        uint32_t st = 0;
        bool changed = false;
        AkIdSink sink;
        sink.buf = stage + threadIdx.x;
        sink.cap = AK_BPE_STAGE;
        sink.stride = AK_BLOCK;
        sink.cnt = 0;
        sink.direct = false;
        sink.gout = A.ids;
        sink.gbase = 0;
        sink.gcap = A.id_cap;
        int64_t row_first = 0, row_last = 0;
        if (sp.s < sp.e)
            ak_bpe_span(A.M, A.T, B.text, B.off, B.n_rows, sp.r_lo, sp.r_hi, sp.s, sp.e, sp.limit, sink, A.id_splits, 0,
                        row_first, row_last, A.pool, changed, st);
        const int cnt = sink.cnt;
        int total;
        const int pre = ak_block_exscan<AK_BLOCK>(cnt, ws, total);
        if (threadIdx.x < 32) {
            long long b = ak_tile_prefix(B.state0, tile, total, (unsigned int*)&B.result[2], AK_ST_SPIN);
            if (threadIdx.x == 0) {
                s_base = b;
                if (tile == B.n_tiles - 1) B.totals[0] = b + total;
            }
        }
        __syncthreads();
        const int64_t obase = s_base + pre;
        if (sp.s < sp.e) {
            if (obase + cnt > A.id_cap) st |= AK_ST_OVERFLOW;
            for (int64_t r = row_first; r < row_last; ++r) A.id_splits[r] += obase;     // span-relative -> global
            if (cnt <= AK_BPE_STAGE) {
                for (int i = 0; i < cnt; ++i)
                    if (obase + i < A.id_cap) A.ids[obase + i] = stage[i * AK_BLOCK + threadIdx.x];
            } else {
                // did not fit the stage (many empty rows or very dense words): walk again straight to global memory
                AkIdSink s2 = sink;
                s2.cnt = 0;
                s2.direct = true;
                s2.gbase = obase;
                uint32_t st2 = 0;
                bool ch2 = false;
                int64_t a, b;
                ak_bpe_span(A.M, A.T, B.text, B.off, B.n_rows, sp.r_lo, sp.r_hi, sp.s, sp.e, sp.limit, s2, nullptr, 0, a, b,
                            A.pool, ch2, st2);
            }
        }
        if (changed) atomicOr(A.changed, 1u);
        ak_raise(B.result, st);
    }
}

// ------------------------------------------------------------------------------------------------
// K2 + K3 fast: grapheme clusters and script runs (ak_seg_fast.cuh), warp tiles.  Two temporary streams (cluster
// ends; run ends + tags), each with per-CTA slices.
// ------------------------------------------------------------------------------------------------
struct AkSfArgs {
    AkBatch B;
    AkTables T;
    uint32_t flags;
    const int64_t* wrow;
    int64_t base0;
    int32_t* tc;                 // temporary streams (sliced per CTA)
    int32_t* tr;
    uint8_t* tt;
    int64_t c_slice, r_slice;
    int32_t* c_total;            // per warp tile
    int64_t* c_toff;
    int32_t* r_total;
    int64_t* r_toff;
    int32_t* c_sums;             // per group of AKW_GROUP warp tiles, and their exclusive prefix
    int64_t* c_sum_base;
    int32_t* r_sums;
    int64_t* r_sum_base;
    AkSegOut o;                  // final outputs
};

__device__ __forceinline__ unsigned long long aks_pack_g(const AkGState& g) {
    return (unsigned long long)g.prev | ((unsigned long long)g.conj << 8) | ((unsigned long long)g.pict << 16) |
           ((unsigned long long)g.ri_odd << 24) | ((unsigned long long)g.prev_m << 32) | ((unsigned long long)g.has_prev << 40);
}
__device__ __forceinline__ AkGState aks_unpack_g(unsigned long long v) {
    AkGState g;
    g.prev = (uint8_t)v; g.conj = (uint8_t)(v >> 8); g.pict = (uint8_t)(v >> 16); g.ri_odd = (uint8_t)(v >> 24);
    g.prev_m = (uint8_t)(v >> 32); g.has_prev = (uint8_t)(v >> 40);
    return g;
}


__global__ void __launch_bounds__(AK_BLOCK, 4) ak_sf_kernel(const AkSfArgs A) {
    __shared__ uint32_t lut[384 + 16];
    __shared__ int32_t cstage[AKS_STAGE * AK_BLOCK];
    __shared__ int32_t rstage[AKS_STAGE * AK_BLOCK];
    __shared__ uint8_t tstage[AKS_STAGE * AK_BLOCK];
    __shared__ unsigned int s_cursor[2];
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    const bool want_c = (A.flags & AK_SEG_CLUSTERS) != 0, want_r = (A.flags & AK_SEG_RUNS) != 0;
    const bool matras = (A.flags & AK_SEG_MATRAS) != 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 384; i += AK_BLOCK)
        lut[i] = i < 128 ? A.T.leaves[((uint32_t)A.T.page_index[0] << 8) | i] : A.T.leaves[((uint32_t)A.T.page_index[9] << 8) | (i - 128)];
    if (tid < 16) lut[384 + tid] = tid < 14 ? aks_pair_row((uint32_t)tid) : 0u;
    if (tid < 2) s_cursor[tid] = 0;
    __syncthreads();
    const int n_wt = akw_n_tiles(B, A.base0);
    const int64_t cslice = (int64_t)blockIdx.x * A.c_slice, rslice = (int64_t)blockIdx.x * A.r_slice;
    for (int wt = blockIdx.x * AKF_WARPS + warp; wt < n_wt; wt += gridDim.x * AKF_WARPS) {
        const int64_t ws = A.base0 + (int64_t)wt * AKF_WARP_BYTES;
        const int64_t r_w0 = A.wrow[wt], r_w1 = A.wrow[wt + 1];
        AkSChunk c;
        const int64_t cs = ws + (int64_t)(lane - 1) * 16;
        akf_load_lane(B, cs, c);
        c.rows = akw_lane_rows(B.off, B.n_rows, r_w0, ws, lane);
        aks_phase_a(A.T, lut, c, matras);
        AkSNeighbor pv;
        pv.g = aks_unpack_g(__shfl_up_sync(0xFFFFFFFFu, aks_pack_g(c.end_g), 1));
        pv.flags = __shfl_up_sync(0xFFFFFFFFu, c.flags, 1);
        pv.end_cur = __shfl_up_sync(0xFFFFFFFFu, c.end_cur, 1);
        const bool real = lane >= 1 && lane <= AKF_REAL;
        const int64_t ss = cs < B.text_begin ? B.text_begin : cs;
        const int64_t se = cs + 16 > B.text_end + 1 ? B.text_end + 1 : cs + 16;
        const bool active = real && ss < se;
        // index of the first row that starts at or after this lane's first position
        int64_t nr = r_w0;
        {
            const int mine = real ? __popc(c.rows & 0xFFFFu) : 0;
            int inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += y;
            }
            nr = r_w0 + (inc - mine);
            if (active) while (nr <= B.n_rows && B.off[nr] < ss) ++nr;      // empty rows share a position
        }
        uint32_t in_cur = AKS_CUR_NONE;
        bool slow = false;
        uint32_t st = 0;
        int64_t row_first = 0, row_last = 0;
        const int64_t rlo = r_w0 > 0 ? r_w0 - 1 : 0, rhi = r_w1 > B.n_rows ? B.n_rows : r_w1;
        AkSegSink sink;
        sink.cbuf = cstage + tid;
        sink.rbuf = rstage + tid;
        sink.tbuf = tstage + tid;
        sink.cap = AKS_STAGE;
        sink.stride = AK_BLOCK;
        sink.cc = sink.rc = 0;
        sink.direct = false;
        sink.gc = sink.gr = nullptr;
        sink.gt = nullptr;
        sink.gccap = sink.grcap = 0;
        if (active) {
            slow = !aks_phase_b(A.T, lut, c, pv, matras, want_c, want_r, in_cur);
            if (slow) {
                AkSegOut o = A.o;
                int64_t scc = 0, src = 0;
                ak_seg_span(A.T, B.text, B.off, B.n_rows, rlo, rhi, ss, se, A.flags, AK_LOOKBACK_LIMIT, false, o, scc, src, st);
                sink.cc = (int)scc;
                sink.rc = (int)src;
            } else {
                aks_lane_emit(c, in_cur, cs, B.off, B.n_rows, nr, want_c, want_r, sink, want_c ? A.o.cluster_splits : nullptr,
                              want_r ? A.o.run_splits : nullptr, row_first, row_last);
            }
        }
        __syncwarp();
        int cinc = sink.cc, rinc = sink.rc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, cinc, d);
            const int z = __shfl_up_sync(0xFFFFFFFFu, rinc, d);
            if (lane >= d) { cinc += y; rinc += z; }
        }
        const int ctot = __shfl_sync(0xFFFFFFFFu, cinc, 31), rtot = __shfl_sync(0xFFFFFFFFu, rinc, 31);
        const int cpre = cinc - sink.cc, rpre = rinc - sink.rc;
        unsigned int ctoff = 0, rtoff = 0;
        if (lane == 0) {
            ctoff = atomicAdd(&s_cursor[0], (unsigned int)ctot);
            rtoff = atomicAdd(&s_cursor[1], (unsigned int)rtot);
        }
        ctoff = __shfl_sync(0xFFFFFFFFu, ctoff, 0);
        rtoff = __shfl_sync(0xFFFFFFFFu, rtoff, 0);
        const bool fits = (int64_t)ctoff + ctot <= A.c_slice && (int64_t)rtoff + rtot <= A.r_slice;
        if (lane == 0) {
            A.c_total[wt] = ctot;
            A.c_toff[wt] = cslice + ctoff;
            A.r_total[wt] = rtot;
            A.r_toff[wt] = rslice + rtoff;
            if (!fits) st |= AK_ST_OVERFLOW;
        }
        if (active) {
            int32_t* cdst = A.tc + cslice + ctoff;
            int32_t* rdst = A.tr + rslice + rtoff;
            uint8_t* tdst = A.tt + rslice + rtoff;
            if (slow) {
                // the walker writes straight into the temporary streams; splits are warp-tile relative like the fast lanes'
                AkSegOut o = A.o;
                o.cluster_ends = cdst;
                o.run_ends = rdst;
                o.run_tags = tdst;
                o.cbase = cpre;
                o.rbase = rpre;
                o.ccap = fits ? ctot : 0;
                o.rcap = fits ? rtot : 0;
                uint32_t st2 = 0;
                int64_t a, b;
                ak_seg_span(A.T, B.text, B.off, B.n_rows, rlo, rhi, ss, se, A.flags, AK_LOOKBACK_LIMIT, true, o, a, b, st2);
            } else {
                for (int64_t r = row_first; r < row_last; ++r) {
                    if (want_c) A.o.cluster_splits[r] += cpre;
                    if (want_r) A.o.run_splits[r] += rpre;
                }
                if (fits) {
                    if (sink.cc <= AKS_STAGE && sink.rc <= AKS_STAGE) {
                        for (int i = 0; i < sink.cc; ++i) cdst[cpre + i] = cstage[i * AK_BLOCK + tid];
                        for (int i = 0; i < sink.rc; ++i) { rdst[rpre + i] = rstage[i * AK_BLOCK + tid]; tdst[rpre + i] = tstage[i * AK_BLOCK + tid]; }
                    } else {
                        AkSegSink s2 = sink;
                        s2.cc = s2.rc = 0;
                        s2.direct = true;
                        s2.gc = cdst + cpre;
                        s2.gr = rdst + rpre;
                        s2.gt = tdst + rpre;
                        s2.gccap = sink.cc;
                        s2.grcap = sink.rc;
                        int64_t a, b;
                        aks_lane_emit(c, in_cur, cs, B.off, B.n_rows, nr, want_c, want_r, s2, nullptr, nullptr, a, b);
                    }
                }
            }
        }
        __syncwarp();
        ak_raise(B.result, st);
    }
}

// ---- K2 + K3 v3: the same outputs from parallel bit streams (ak_seg3.cuh): 32 bytes per lane, a warp covers two
// 480-byte warp tiles (lanes 1-15 and 16-30), so the bookkeeping per warp tile -- totals, temporary-stream offsets,
// tile-relative splits -- and with it the sums / scan / copy kernels stay as they are.  Counts are popcounts of the
// event masks, known before anything is written: no shared-memory staging, the events go straight to the lane's
// place in the temporary stream.
#define AKS3_THREADS 128
#ifndef AKS3_MINB
#define AKS3_MINB 8
#endif
__global__ void __launch_bounds__(AKS3_THREADS, AKS3_MINB) ak_sf3_kernel(const AkSfArgs A) {
    __shared__ unsigned int s_cursor[2];
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    const bool want_c = (A.flags & AK_SEG_CLUSTERS) != 0, want_r = (A.flags & AK_SEG_RUNS) != 0;
    const bool matras = (A.flags & AK_SEG_MATRAS) != 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2) s_cursor[tid] = 0;
    __syncthreads();
    const int n_wt = akw_n_tiles(B, A.base0);
    const int n_w3 = (n_wt + 1) >> 1;
    const int64_t tb = B.text_begin, te = B.text_end;
    const int64_t cslice = (int64_t)blockIdx.x * A.c_slice, rslice = (int64_t)blockIdx.x * A.r_slice;
    for (int w3 = blockIdx.x * (AKS3_THREADS / 32) + warp; w3 < n_w3; w3 += gridDim.x * (AKS3_THREADS / 32)) {
        const int wt0 = 2 * w3;
        const bool two = wt0 + 1 < n_wt;
        const int64_t ws = A.base0 + (int64_t)wt0 * AKF_WARP_BYTES;
        const int64_t r_w0 = A.wrow[wt0], r_w2 = A.wrow[two ? wt0 + 2 : wt0 + 1];
        const int64_t cs = ws + (int64_t)(lane - 1) * 32;
        AkS3Lane L;
        {
            uint32_t x[8];
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
            L.rows = akn3_lane_rows(B.off, B.n_rows, r_w0, ws, lane);
            aks3_phase1(x, L);
        }
        uint32_t dn1n = __shfl_down_sync(0xFFFFFFFFu, L.dn1, 1);
        if (lane == 31) dn1n = 0;
        aks3_phase2(L, dn1n);
        if (L.FOR) aks3_foreign(A.T, B.text, cs, te, L);
        aks3_summary(L);
        const uint32_t up2p = __shfl_up_sync(0xFFFFFFFFu, L.up2, 1);
        const bool real = lane >= 1 && lane <= 30;
        const int64_t ss = cs < tb ? tb : cs;
        const int64_t se = cs + 32 > te + 1 ? te + 1 : cs + 32;
        const bool active = real && ss < se && (two || lane <= 15);
        // index of the first row that starts at or after this lane's first position
        int64_t nr;
        {
            const int mine = real ? __popc(L.rows) : 0;
            int inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += y;
            }
            nr = r_w0 + (inc - mine);
            if (active) while (nr <= B.n_rows && B.off[nr] < ss) ++nr;
        }
        const uint32_t tb_bit = (tb >= cs && tb < cs + 32) ? 1u << (int)(tb - cs) : 0u;
        const uint32_t rows_ev = L.rows & ~tb_bit;
        bool slow = false;
        uint32_t st = 0;
        int cc = 0, rc = 0;
        const int64_t rlo = r_w0 > 0 ? r_w0 - 1 : 0, rhi = r_w2 > B.n_rows ? B.n_rows : r_w2;
        if (active) {
            slow = !aks3_phase3(L, up2p, tb_bit, matras, want_c, want_r);
            if (slow) {
                AkSegOut o = A.o;
                int64_t scc = 0, src = 0;
                ak_seg_span(A.T, B.text, B.off, B.n_rows, rlo, rhi, ss, se, A.flags, AK_LOOKBACK_LIMIT, false, o, scc, src, st);
                cc = (int)scc;
                rc = (int)src;
            } else {
                const int nre = __popc(rows_ev);
                if (want_c) cc = __popc(L.brk) + nre;
                if (want_r) rc = __popc(L.rchg) + nre;
            }
        }
        // one scan for both counts (a lane has at most 33 events per stream)
        int inc2 = cc | (rc << 16);
        const int mine2 = inc2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, inc2, d);
            if (lane >= d) inc2 += y;
        }
        const int tot2 = __shfl_sync(0xFFFFFFFFu, inc2, 31), half2 = __shfl_sync(0xFFFFFFFFu, inc2, 15);
        const int ctot = tot2 & 0xFFFF, rtot = tot2 >> 16, chalf = half2 & 0xFFFF, rhalf = half2 >> 16;
        const int cpre = (inc2 - mine2) & 0xFFFF, rpre = (inc2 - mine2) >> 16;
        unsigned int ctoff = 0, rtoff = 0;
        if (lane == 0) {
            ctoff = atomicAdd(&s_cursor[0], (unsigned int)ctot);
            rtoff = atomicAdd(&s_cursor[1], (unsigned int)rtot);
        }
        ctoff = __shfl_sync(0xFFFFFFFFu, ctoff, 0);
        rtoff = __shfl_sync(0xFFFFFFFFu, rtoff, 0);
        const bool fits = (int64_t)ctoff + ctot <= A.c_slice && (int64_t)rtoff + rtot <= A.r_slice;
        if (lane == 0) {
            A.c_total[wt0] = chalf;
            A.c_toff[wt0] = cslice + ctoff;
            A.r_total[wt0] = rhalf;
            A.r_toff[wt0] = rslice + rtoff;
            if (two) {
                A.c_total[wt0 + 1] = ctot - chalf;
                A.c_toff[wt0 + 1] = cslice + ctoff + chalf;
                A.r_total[wt0 + 1] = rtot - rhalf;
                A.r_toff[wt0 + 1] = rslice + rtoff + rhalf;
            }
            if (!fits) st |= AK_ST_OVERFLOW;
        }
        if (active) {
            const bool second = lane > 15;
            const int cpre_t = second ? cpre - chalf : cpre, rpre_t = second ? rpre - rhalf : rpre;      // tile-relative
            int32_t* cdst = A.tc + cslice + ctoff;
            int32_t* rdst = A.tr + rslice + rtoff;
            uint8_t* tdst = A.tt + rslice + rtoff;
            if (slow) {
                AkSegOut o = A.o;
                o.cluster_ends = cdst + (second ? chalf : 0);
                o.run_ends = rdst + (second ? rhalf : 0);
                o.run_tags = tdst + (second ? rhalf : 0);
                o.cbase = cpre_t;
                o.rbase = rpre_t;
                o.ccap = fits ? (second ? ctot - chalf : chalf) : 0;
                o.rcap = fits ? (second ? rtot - rhalf : rhalf) : 0;
                uint32_t st2 = 0;
                int64_t a, b;
                ak_seg_span(A.T, B.text, B.off, B.n_rows, rlo, rhi, ss, se, A.flags, AK_LOOKBACK_LIMIT, true, o, a, b, st2);
            } else {
                const uint32_t mc = want_c ? (L.brk | rows_ev) : 0u, mr = want_r ? (L.rchg | rows_ev) : 0u;
                if (fits) {
                    const int64_t rs_in = nr > 0 ? B.off[nr - 1] : B.off[0];
                    if (want_c) aks3_emit(L, mc, cs, rs_in, cdst + cpre, nullptr);
                    if (want_r) aks3_emit(L, mr, cs, rs_in, rdst + rpre, tdst + rpre);
                }
                if (L.rows)
                    aks3_splits(L, mc, mr, cs, B.off, B.n_rows, nr, cpre_t, rpre_t, want_c ? A.o.cluster_splits : nullptr,
                                want_r ? A.o.run_splits : nullptr);
            }
        }
        ak_raise(B.result, st);
    }
}

// Flat copy of one warp's 32 consecutive warp-tile blocks from the temporary stream to their (contiguous) final range:
// lane k moves elements k, k + 32, ... of the whole range, four loads in flight; the tile an element belongs to comes
// from the exclusive prefix in shared memory (s_excl[0..32], s_delta[j] = block start in temp - exclusive prefix).
template <class T>
__device__ __forceinline__ void akw_flat_copy(const T* temp, T* out, int64_t dst0, int W, const int* s_excl, const long long* s_delta,
                                              int lane) {
    int j = 0;
    for (int k = lane; k < W; k += 128) {
        long long sidx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int kk = k + 32 * u;
            if (kk < W) {
                while (kk >= s_excl[j + 1]) ++j;
                sidx[u] = s_delta[j] + kk;
            } else sidx[u] = -1;
        }
        T v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (sidx[u] >= 0) v[u] = temp[sidx[u]];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (sidx[u] >= 0) out[dst0 + k + 32 * u] = v[u];
    }
}

__global__ void __launch_bounds__(AKW_GROUP) ak_sf_copy_kernel(const AkSfArgs A) {
    __shared__ int ws[33];
    __shared__ int s_excl[AKW_GROUP / 32][33];
    __shared__ long long s_delta[AKW_GROUP / 32][32];
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    const bool want_c = (A.flags & AK_SEG_CLUSTERS) != 0, want_r = (A.flags & AK_SEG_RUNS) != 0;
    const int n_wt = akw_n_tiles(B, A.base0);
    const int n_groups = (n_wt + AKW_GROUP - 1) / AKW_GROUP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int gidx = blockIdx.x; gidx < n_groups; gidx += gridDim.x) {
        const int t = gidx * AKW_GROUP + tid;
        const int64_t r0 = t < n_wt ? A.wrow[t] : 0, r1 = t < n_wt ? A.wrow[t + 1] : 0;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 0 ? !want_c : !want_r) continue;
            const int32_t* totals = pass == 0 ? A.c_total : A.r_total;
            const int64_t* toffs = pass == 0 ? A.c_toff : A.r_toff;
            int64_t* splits = pass == 0 ? A.o.cluster_splits : A.o.run_splits;
            const int mine = t < n_wt ? totals[t] : 0;
            int total;
            const int pre = ak_block_exscan<AKW_GROUP>(mine, ws, total);
            const int64_t dst = (pass == 0 ? A.c_sum_base[gidx] : A.r_sum_base[gidx]) + pre;
            for (int64_t r = r0; r < r1 && r <= B.n_rows; ++r) splits[r] += dst;       // rows that start in my warp tile
            const int pre_w = __shfl_sync(0xFFFFFFFFu, pre, 0);
            __syncwarp();
            s_excl[warp][lane] = pre - pre_w;
            s_delta[warp][lane] = (t < n_wt ? toffs[t] : 0) - (long long)(pre - pre_w);
            const int W = __shfl_sync(0xFFFFFFFFu, pre + mine, 31) - pre_w;
            if (lane == 0) s_excl[warp][32] = 0x7FFFFFFF;
            __syncwarp();
            const int64_t dst0 = __shfl_sync(0xFFFFFFFFu, dst, 0);
            if (dst0 + W > (pass == 0 ? A.o.ccap : A.o.rcap)) { if (lane == 0 && W > 0) ak_raise(B.result, AK_ST_OVERFLOW); }
            else if (pass == 0) akw_flat_copy<int32_t>(A.tc, A.o.cluster_ends, dst0, W, s_excl[warp], s_delta[warp], lane);
            else {
                akw_flat_copy<int32_t>(A.tr, A.o.run_ends, dst0, W, s_excl[warp], s_delta[warp], lane);
                akw_flat_copy<uint8_t>(A.tt, A.o.run_tags, dst0, W, s_excl[warp], s_delta[warp], lane);
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4a fast: BPE through the word cache (ak_bpe_fast.cuh), warp tiles
// ------------------------------------------------------------------------------------------------
struct AkBfArgs {
    AkBatch B;
    AkTables T;
    AkBpeDev M;
    AkWordCache C;
    AkPool pool;
    const int64_t* wrow;         // [n_wt + 1]
    int64_t base0;
    int32_t* temp;               // temporary id stream, one private slice per CTA of the encode kernel
    int64_t slice_cap;
    int32_t* wt_total;           // [n_wt]
    int64_t* wt_toff;            // [n_wt] where the warp tile's block starts in temp
    int32_t* sums;               // [groups] and their exclusive prefix
    int64_t* sum_base;
    int32_t* ids;
    int64_t id_cap;
    int64_t* id_splits;
    unsigned int* changed;
};

// cold parts of the event loop (a word that is not in the cache, or cannot be cached), out of line on purpose
__device__ __noinline__ uint32_t akb_event_slow(const AkBfArgs& A, const AkBatch& B, int64_t p, uint32_t len, uint32_t kc,
                                                unsigned long long h, unsigned long long want, long long slot, bool cacheable,
                                                int64_t r_lo, int64_t r_hi, uint32_t& st) {
    int64_t we = p + len;
    if (len == 0xFFFFFu)       // clamped in the event record: find the real end again
        we = akb_word_end(A.T, B.text, p, 0, kc, 0u, 0u, B.off, B.n_rows, r_lo, r_hi);
    if (cacheable) {
        int32_t tmp[AKW_MAXTOK + 1];
        AkIdSink local;
        local.buf = tmp; local.cap = AKW_MAXTOK + 1; local.stride = 1; local.cnt = 0; local.direct = false;
        local.gout = nullptr; local.gbase = 0; local.gcap = 0;
        ak_bpe_word(A.M, A.T, B.text, p, we, kc, local, A.pool, st);
        const int n = local.cnt;
        if (n <= AKW_MAXTOK && slot >= 0) {
            akw_insert(A.C, slot, want, B.text, p, len, tmp, n);
            long long s2;
            const long long hit = akw_find(A.C, h, want, B.text, p, len, &s2);      // ours, or the same word by another lane
            if (hit >= 0) {
                const int m = (int)((akw_ld(A.C.e + (unsigned long long)hit * AKW_ENTRY) & AKW_NTOK_MASK) >> 3);
                return ((uint32_t)hit << 5) | (uint32_t)m;
            }
        }
        return 0x80000000u | (uint32_t)n;
    }
    AkIdSink cntsink;
    cntsink.buf = nullptr; cntsink.cap = 0; cntsink.stride = 1; cntsink.cnt = 0; cntsink.direct = false;
    cntsink.gout = nullptr; cntsink.gbase = 0; cntsink.gcap = 0;
    ak_bpe_word(A.M, A.T, B.text, p, we, kc, cntsink, A.pool, st);
    return 0x80000000u | (uint32_t)cntsink.cnt;
}

__device__ __noinline__ void akb_event_write_direct(const AkBfArgs& A, const AkBatch& B, int64_t p, uint32_t len, uint32_t kc,
                                                    int64_t r_lo, int64_t r_hi, int32_t* dst, int64_t n) {
    int64_t we = p + len;
    if (len == 0xFFFFFu) we = akb_word_end(A.T, B.text, p, 0, kc, 0u, 0u, B.off, B.n_rows, r_lo, r_hi);
    AkIdSink ds;
    ds.buf = nullptr; ds.cap = 0; ds.stride = 1; ds.cnt = 0; ds.direct = true;
    ds.gout = dst; ds.gbase = 0; ds.gcap = n;
    uint32_t st2 = 0;
    ak_bpe_word(A.M, A.T, B.text, p, we, kc, ds, A.pool, st2);
}

// more events than the list holds, or more than 65535 ids in one warp tile (hundreds of one-byte words / thousands of
// empty rows in 480 bytes): lane by lane, counted then written straight to the temporary stream.  Cold.
__device__ __noinline__ void akb_tile_fallback(const AkBfArgs& A, const AkBatch& B, const AkBLaneCtx& X, const AkBChunk& c,
                                               uint32_t next_bnd, int64_t cs, bool active, int lane, int wt, int64_t slice,
                                               int64_t r_w0, int rows_before, unsigned int* s_cursor_p, uint32_t& st) {
    int total = 0;
    int64_t nr_hint = -1;
    if (active && c.rows) {
        int64_t g = r_w0 + (rows_before);
        const int64_t p = cs + (__ffs(c.rows) - 1);
        while (g < B.n_rows && B.off[g] < p) ++g;
        nr_hint = g;
    }
    AkIdSink sink;
    sink.buf = nullptr; sink.cap = 0; sink.stride = 1; sink.cnt = 0; sink.direct = false;
    sink.gout = A.temp; sink.gbase = 0; sink.gcap = 0;
    int64_t row_first = 0, row_last = 0;
    if (active) akb_lane_emit(X, c, next_bnd, cs, sink, A.id_splits, row_first, row_last, st, nr_hint);
    const int cnt = sink.cnt;
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += y;
    }
    total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    const int pre = inc - cnt;
    unsigned int toff = 0;
    if (lane == 0) toff = atomicAdd(s_cursor_p, (unsigned int)total);
    toff = __shfl_sync(0xFFFFFFFFu, toff, 0);
    const bool fits = (int64_t)toff + total <= A.slice_cap;
    if (lane == 0) {
        A.wt_total[wt] = total;
        A.wt_toff[wt] = slice + toff;
        if (!fits) st |= AK_ST_OVERFLOW;
    }
    if (active) {
        for (int64_t r = row_first; r < row_last; ++r) A.id_splits[r] += pre;
        if (fits) {
            AkIdSink s2 = sink;
            s2.cnt = 0;
            s2.direct = true;
            s2.gout = A.temp + slice + toff + pre;
            s2.gbase = 0;
            s2.gcap = cnt;
            uint32_t st2 = 0;
            int64_t a, b;
            akb_lane_emit(X, c, next_bnd, cs, s2, nullptr, a, b, st2, nr_hint);
        }
    }
}

// one 480-byte warp tile of the v2 (16 bytes per lane) encoder: classification, event list, cache look-ups, ids to the
// CTA's slice of the temporary stream.  Also the fallback of the bit-parallel kernel for warp tiles with too many events.
__device__ __noinline__ void akb_v2_warp_tile(const AkBfArgs& A, const AkBatch& B, int wt, int lane, const uint32_t* lut,
                                              uint32_t* ev_w, uint32_t* res_w, uint16_t* eoff_w, unsigned int* s_cursor, int64_t slice) {
    const int64_t ws = A.base0 + (int64_t)wt * AKF_WARP_BYTES;
    const int64_t r_w0 = A.wrow[wt], r_w1 = A.wrow[wt + 1];
    AkBChunk c;
    const int64_t cs = ws + (int64_t)(lane - 1) * 16;
    akf_load_lane(B, cs, c);
    c.rows = akw_lane_rows(B.off, B.n_rows, r_w0, ws, lane);
    akb_phase_a(A.T, lut, c);
    {
        uint32_t pw = __shfl_up_sync(0xFFFFFFFFu, c.last_w, 1);
        uint32_t pk = __shfl_up_sync(0xFFFFFFFFu, c.last_cls, 1);
        if (lane == 0) {
            pw = AKF_NONE;
            pk = 2;
            if (c.first_pos < 32u && cs > B.text_begin) {
                int64_t q = cs - 1;
                int k = 0;
                while (q > B.text_begin && k < 3 && (B.text[q] & 0xC0u) == 0x80u) { --q; ++k; }
                int len;
                pw = akf_props(A.T, lut, ak_decode(B.text, q, B.text_end, len));
                pk = AK_HFCLASS(pw);
            }
        }
        akb_resolve_first(c, pw, pk);
    }
    // word boundaries of the next chunk (bits 0..15) and of the one after it (bits 16..31; unknown for lane 30)
    uint32_t next_bnd = __shfl_down_sync(0xFFFFFFFFu, c.bnd, 1) & 0xFFFFu;
    {
        const uint32_t n2 = __shfl_down_sync(0xFFFFFFFFu, c.bnd, 2) & 0xFFFFu;
        if (lane < 30) next_bnd |= n2 << 16;
    }
    const bool real = lane >= 1 && lane <= AKF_REAL;
    const int64_t ss = cs < B.text_begin ? B.text_begin : cs;
    const int64_t se = cs + 16 > B.text_end + 1 ? B.text_end + 1 : cs + 16;
    const bool active = real && ss < se;
    AkBLaneCtx X;
    X.M = &A.M;
    X.T = &A.T;
    X.C = &A.C;
    X.text = B.text;
    X.off = B.off;
    X.n_rows = B.n_rows;
    X.r_lo = r_w0 > 0 ? r_w0 - 1 : 0;
    X.r_hi = r_w1 > B.n_rows ? B.n_rows : r_w1;
    X.pool = &A.pool;
    uint32_t st = 0;
    if (active) {
        if (c.flags & AKB_ALPHABET) st |= AK_ST_ALPHABET;
        if ((c.flags & AKF_TROUBLE) && akb_chunk_changes(X, c, cs, AK_LOOKBACK_LIMIT, st)) atomicOr(A.changed, 1u);
    }
    // ---- the warp tile's EVENT LIST: row starts and word starts in position order, so that the words can be
    // encoded one per lane, 32 at a time, whatever chunk they came from
    uint32_t wstart = 0;
    if (active) {
        const uint32_t hi = (c.cls >> 1) & 0x55555555u & ~c.cls;      // bit 2i set <=> class at byte i is 2 (space)
        uint32_t x = hi;
        x = (x | (x >> 1)) & 0x33333333u;
        x = (x | (x >> 2)) & 0x0F0F0F0Fu;
        x = (x | (x >> 4)) & 0x00FF00FFu;
        x = (x | (x >> 8)) & 0x0000FFFFu;
        wstart = c.bnd & c.lead & ~x;
    }
    const uint32_t rowsm = active ? (c.rows & 0xFFFFu) : 0u;
    const int n_row_ev = __popc(rowsm), n_ev = n_row_ev + __popc(wstart);
    int e_inc = n_ev, r_inc = n_row_ev;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, e_inc, d);
        const int z = __shfl_up_sync(0xFFFFFFFFu, r_inc, d);
        if (lane >= d) { e_inc += y; r_inc += z; }
    }
    const int E = __shfl_sync(0xFFFFFFFFu, e_inc, 31);
    uint32_t* ev = ev_w;
    uint32_t* res = res_w;
    uint16_t* eoff = eoff_w;
    bool use_list = E <= AKB_EVCAP;
    int total = 0;
    if (use_list) {
        {
            int k = e_inc - n_ev, rord = r_inc - n_row_ev;
            uint32_t m = rowsm | wstart;
            while (m) {
                const int i = __ffs(m) - 1;
                m &= m - 1u;
                const uint32_t pos = (uint32_t)(cs + i - ws);
                if ((rowsm >> i) & 1u) ev[k++] = pos | (1u << 9) | ((uint32_t)rord++ << 12);
                if ((wstart >> i) & 1u) {
                    const uint32_t kc = (c.cls >> (2 * i)) & 3u;
                    const int64_t e = akb_word_end(A.T, B.text, cs, i, kc, c.bnd, next_bnd, B.off, B.n_rows, X.r_lo, X.r_hi);
                    int64_t len = e - (cs + i);
                    if (len > 0xFFFFF) len = 0xFFFFF;
                    ev[k++] = pos | (kc << 10) | ((uint32_t)len << 12);
                }
            }
        }
        __syncwarp();
        // ---- pass 1: tokens per event (cache lookup; a miss runs the merge loop and publishes the word)
        int running = 0;
        for (int base = 0; base < E; base += 32) {
            const int e = base + lane;
            int n = 0;
            uint32_t rr = 0;
            if (e < E) {
                const uint32_t v = ev[e];
                const int64_t p = ws + (v & 511u);
                if (v & (1u << 9)) {
                    int64_t g = r_w0 + (v >> 12);
                    while (g < B.n_rows && B.off[g] < p) ++g;
                    while (g <= B.n_rows && B.off[g] == p) {
                        if (g > 0 && A.M.eos >= 0) ++n;
                        if (g < B.n_rows && A.M.bos >= 0) ++n;
                        ++g;
                    }
                    rr = 0x40000000u | (uint32_t)n;
                } else {
                    const uint32_t kc = (v >> 10) & 3u;
                    const uint32_t len = v >> 12;
                    // hot path: hash, probe, compare -- everything else lives in akb_event_slow (kept out of line so
                    // that this loop stays small in the instruction cache)
                    long long hit = -1, slot = -1;
                    unsigned long long h = 0, want = 0;
                    const bool cacheable = len <= AKW_MAXLEN && A.C.e != nullptr;
                    if (cacheable) {
                        h = akw_hash(B.text, p, len);
                        want = akw_want(h, len);
                        hit = akw_find(A.C, h, want, B.text, p, len, &slot);
                    }
                    if (hit >= 0) {
                        n = (int)((akw_ld(A.C.e + (unsigned long long)hit * AKW_ENTRY) & AKW_NTOK_MASK) >> 3);
                        rr = ((uint32_t)hit << 5) | (uint32_t)n;
                    } else {
                        rr = akb_event_slow(A, B, p, len, kc, h, want, slot, cacheable, X.r_lo, X.r_hi, st);
                        n = (rr & 0x80000000u) ? (int)(rr & 0x3FFFFFFFu) : (int)(rr & 31u);
                    }
                }
            }
            int inc = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += y;
            }
            if (e < E) {
                res[e] = rr;
                eoff[e] = (uint16_t)(running + inc - n);
            }
            running += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        total = running;
        if (total >= 65536) use_list = false;      // offsets are 16-bit (thousands of empty rows at one position)
    }
    if (use_list) {
        unsigned int toff = 0;
        if (lane == 0) toff = atomicAdd(s_cursor, (unsigned int)total);
        toff = __shfl_sync(0xFFFFFFFFu, toff, 0);
        const bool fits = (int64_t)toff + total <= A.slice_cap;
        if (lane == 0) {
            A.wt_total[wt] = total;
            A.wt_toff[wt] = slice + toff;
            if (!fits) st |= AK_ST_OVERFLOW;
        }
        __syncwarp();
        // ---- pass 2: write the ids (from the cache entries) and the row splits (warp-tile relative)
        int32_t* tbase = A.temp + slice + toff;
        for (int base = 0; base < E; base += 32) {
            const int e = base + lane;
            if (e >= E) continue;
            const uint32_t v = ev[e], rr = res[e];
            const int o = eoff[e];
            const int64_t p = ws + (v & 511u);
            if (v & (1u << 9)) {
                int64_t g = r_w0 + (v >> 12);
                while (g < B.n_rows && B.off[g] < p) ++g;
                int k = o;
                while (g <= B.n_rows && B.off[g] == p) {
                    if (g > 0 && A.M.eos >= 0) { if (fits) tbase[k] = A.M.eos; ++k; }
                    A.id_splits[g] = k;
                    if (g < B.n_rows && A.M.bos >= 0) { if (fits) tbase[k] = A.M.bos; ++k; }
                    ++g;
                }
            } else if (fits) {
                const int n = (int)(rr & 31u);
                if (!(rr & 0x80000000u)) {
                    const unsigned long long* en = A.C.e + (unsigned long long)((rr >> 5) & 0x3FFFFu) * AKW_ENTRY;
#pragma unroll 1
                    for (int i = 0; i < n; i += 2) {
                        const unsigned long long q = akw_ld(en + 8 + (i >> 1));
                        tbase[o + i] = (int32_t)(uint32_t)q;
                        if (i + 1 < n) tbase[o + i + 1] = (int32_t)(uint32_t)(q >> 32);
                    }
                } else {
                    akb_event_write_direct(A, B, p, v >> 12, (v >> 10) & 3u, X.r_lo, X.r_hi, tbase + o, (int64_t)(rr & 0x3FFFFFFFu));
                }
            }
        }
    } else {
        akb_tile_fallback(A, B, X, c, next_bnd, cs, active, lane, wt, slice, r_w0, r_inc - n_row_ev, s_cursor, st);
    }
    __syncwarp();
    ak_raise(B.result, st);
}

#ifndef AKB_MINB
#define AKB_MINB 4
#endif
__global__ void __launch_bounds__(AK_BLOCK, AKB_MINB) ak_bf_encode_kernel(const AkBfArgs A) {
    __shared__ uint32_t lut[384];
    __shared__ uint32_t s_ev[AKB_EVCAP * AKF_WARPS];       // per warp: the tile's events (row starts, word starts)
    __shared__ uint32_t s_res[AKB_EVCAP * AKF_WARPS];      // per event: cache entry + token count
    __shared__ uint16_t s_eoff[AKB_EVCAP * AKF_WARPS];     // per event: token offset inside the warp tile's block
    __shared__ unsigned int s_cursor;
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 384; i += AK_BLOCK)
        lut[i] = i < 128 ? A.T.leaves[((uint32_t)A.T.page_index[0] << 8) | i] : A.T.leaves[((uint32_t)A.T.page_index[9] << 8) | (i - 128)];
    if (tid == 0) s_cursor = 0;
    __syncthreads();
    const int n_wt = akw_n_tiles(B, A.base0);
    const int64_t slice = (int64_t)blockIdx.x * A.slice_cap;
    for (int wt0 = blockIdx.x * AKF_WARPS; wt0 < n_wt; wt0 += gridDim.x * AKF_WARPS) {
        const int wt = wt0 + warp;
        if (wt >= n_wt) continue;
        akb_v2_warp_tile(A, B, wt, lane, lut, s_ev + warp * AKB_EVCAP, s_res + warp * AKB_EVCAP, s_eoff + warp * AKB_EVCAP, &s_cursor, slice);
    }
}

// ---- K4a v3: the same encoder behind a bit-parallel front end (ak_bpe3.cuh).  32 bytes per lane, a warp covers two
// 480-byte warp tiles (lanes 1-15 / 16-30) with ONE event list, so the per-tile bookkeeping and the sums / scan / copy
// kernels are shared with v2.  Event record: position in the 960 bytes (10 bits) | row flag (bit 10) | class (bits 11-12) |
// word length or row ordinal (from bit 13).
#define AKB3_THREADS 128
#define AKB3_WARPS (AKB3_THREADS / 32)
#ifndef AKB3_EVCAP
#define AKB3_EVCAP 512       // events per 960 bytes kept in shared memory (a denser tile takes the v2 routine)
#endif
#ifndef AKB3_MINB
#define AKB3_MINB 8         // measured: 5 -> 3.28 ms, 6 -> 3.26, 8 (64 registers, 512-event lists) -> 3.09 per 256 MiB
#endif

// exact NFC check of the troubled code points (cold): does NFC change the text?
__device__ __noinline__ bool akb3_changes(const AkTables& T, const uint8_t* text, const int64_t* off, int64_t n_rows, int64_t r_lo,
                                          uint32_t trb, int64_t cs, uint32_t& status) {
    bool changed = false;
    int64_t checked_until = -1;
    while (trb) {
        const int i = __ffs(trb) - 1;
        trb &= trb - 1u;
        const int64_t p = cs + i;
        if (p < checked_until) continue;
        const int64_t r = ak_row_lower_bound(off, r_lo, n_rows, p + 1);
        const int64_t rs = off[r - 1], re = off[r];
        if (ak_segment_changes(T, text, p, rs, re, AK_LOOKBACK_LIMIT, &checked_until, status)) changed = true;
    }
    return changed;
}

// end of a word of class k with no boundary before `from` (cold: words longer than 64 bytes)
__device__ __noinline__ int64_t akb3_scan_end(const AkTables& T, const uint8_t* t, int64_t wpos, int64_t from, uint32_t k,
                                              const int64_t* off, int64_t n_rows, int64_t r_lo, int64_t r_hi) {
    int64_t er = ak_row_lower_bound(off, r_lo, r_hi, wpos + 1);
    if (off[er] < wpos + 1) er = ak_row_lower_bound(off, r_hi, n_rows, wpos + 1);
    const int64_t re = off[er];
    int64_t q = from;
    if (q > re) q = re;
    while (q < re && (t[q] & 0xC0u) == 0x80u) ++q;
    while (q < re) {
        int len;
        const uint32_t cp = ak_decode(t, q, re, len);
        if (AK_HFCLASS(ak_props(T, cp)) != k) break;
        q += len;
    }
    return q;
}

__global__ void __launch_bounds__(AKB3_THREADS, AKB3_MINB) ak_bf3_encode_kernel(const AkBfArgs A) {
    __shared__ uint32_t lut[384];                              // only the fallback (akb_v2_warp_tile) reads it
    __shared__ uint32_t s_ev[AKB3_EVCAP * AKB3_WARPS];
    __shared__ uint32_t s_res[AKB3_EVCAP * AKB3_WARPS];
    __shared__ uint16_t s_eoff[AKB3_EVCAP * AKB3_WARPS];
    __shared__ unsigned int s_cursor;
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 384; i += AKB3_THREADS)
        lut[i] = i < 128 ? A.T.leaves[((uint32_t)A.T.page_index[0] << 8) | i] : A.T.leaves[((uint32_t)A.T.page_index[9] << 8) | (i - 128)];
    if (tid == 0) s_cursor = 0;
    __syncthreads();
    const int n_wt = akw_n_tiles(B, A.base0);
    const int n_w3 = (n_wt + 1) >> 1;
    const int64_t tb = B.text_begin, te = B.text_end;
    const int64_t slice = (int64_t)blockIdx.x * A.slice_cap;
    uint32_t* ev = s_ev + warp * AKB3_EVCAP;
    uint32_t* res = s_res + warp * AKB3_EVCAP;
    uint16_t* eoff = s_eoff + warp * AKB3_EVCAP;
    for (int w3 = blockIdx.x * AKB3_WARPS + warp; w3 < n_w3; w3 += gridDim.x * AKB3_WARPS) {
        const int wt0 = 2 * w3;
        const bool two = wt0 + 1 < n_wt;
        const int64_t ws = A.base0 + (int64_t)wt0 * AKF_WARP_BYTES;
        const int64_t r_w0 = A.wrow[wt0], r_w2 = A.wrow[two ? wt0 + 2 : wt0 + 1];
        const int64_t cs = ws + (int64_t)(lane - 1) * 32;
        const int64_t r_lo = r_w0 > 0 ? r_w0 - 1 : 0, r_hi = r_w2 > B.n_rows ? B.n_rows : r_w2;
        AkB3Lane L;
        {
            uint32_t x[8];
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
            L.rows = akn3_lane_rows(B.off, B.n_rows, r_w0, ws, lane);
            akb3_phase1(x, L);
        }
        uint32_t dn1n = __shfl_down_sync(0xFFFFFFFFu, L.dn1, 1);
        if (lane == 31) dn1n = 0;
        akb3_phase2(L, dn1n);
        if (L.FOR) akb3_foreign(A.T, B.text, cs, te, L);
        akb3_summary(L);
        const uint32_t up2p = __shfl_up_sync(0xFFFFFFFFu, L.up2, 1);
        akb3_phase3(L, up2p);
        const bool real = lane >= 1 && lane <= 30;
        const int64_t ss = cs < tb ? tb : cs;
        const int64_t se = cs + 32 > te + 1 ? te + 1 : cs + 32;
        const bool active = real && ss < se && (two || lane <= 15);
        uint32_t st = 0;
        if (active) {
            if (L.flags & 1u) st |= AK_ST_ALPHABET;
            if (L.trb && akb3_changes(A.T, B.text, B.off, B.n_rows, r_lo, L.trb, cs, st)) atomicOr(A.changed, 1u);
        }
        // boundaries of the next two lanes (the right halo lane classified its last two bytes without look-ahead)
        const uint32_t bsend = lane == 31 ? (L.bnd & 0x3FFFFFFFu) : L.bnd;
        const uint32_t nb1 = __shfl_down_sync(0xFFFFFFFFu, bsend, 1);
        uint32_t nb2 = __shfl_down_sync(0xFFFFFFFFu, bsend, 2);
        if (lane >= 30) nb2 = 0;
        const uint32_t wstart = active ? L.wstart : 0u;
        const uint32_t rowsm = active ? L.rows : 0u;
        const int n_row_ev = __popc(rowsm), n_ev = n_row_ev + __popc(wstart);
        int e_inc = n_ev, r_inc = n_row_ev;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, e_inc, d);
            const int z = __shfl_up_sync(0xFFFFFFFFu, r_inc, d);
            if (lane >= d) { e_inc += y; r_inc += z; }
        }
        const int E = __shfl_sync(0xFFFFFFFFu, e_inc, 31);
        const int E480 = __shfl_sync(0xFFFFFFFFu, e_inc, 15);          // events of the first warp tile
        bool use_list = E <= AKB3_EVCAP;
        int total = 0;
        if (use_list) {
            {
                int k = e_inc - n_ev, rord = r_inc - n_row_ev;
                uint32_t m = rowsm | wstart;
                while (m) {
                    const int i = __ffs(m) - 1;
                    m &= m - 1u;
                    const uint32_t pos = (uint32_t)(cs + i - ws);
                    if ((rowsm >> i) & 1u) {
                        // the rows that start here (several when rows are empty): their </s> <s> count, and the index
                        // of the first one, are settled now so that the two passes below never search the offsets
                        int64_t g = r_w0 + rord++;
                        const int64_t p = cs + i;
                        while (g < B.n_rows && B.off[g] < p) ++g;
                        const int64_t g0 = g;
                        int n = 0;
                        while (g <= B.n_rows && B.off[g] == p) {
                            if (g > 0 && A.M.eos >= 0) ++n;
                            if (g < B.n_rows && A.M.bos >= 0) ++n;
                            ++g;
                        }
                        res[k] = 0x40000000u | (uint32_t)n;
                        ev[k++] = pos | (1u << 10) | ((uint32_t)(g0 - r_w0) << 13);
                    }
                    if ((wstart >> i) & 1u) {
                        const uint32_t kc = (L.CW >> i) & 1u;
                        const uint32_t above = L.bnd & ~((2u << i) - 1u);
                        int64_t len;
                        if (above) len = (__ffs(above) - 1) - i;
                        else if (nb1) len = 32 - i + (__ffs(nb1) - 1);
                        else if (nb2) len = 64 - i + (__ffs(nb2) - 1);
                        else    // no boundary in what the warp knows: 96 bytes, less where the right halo lane's last two bytes are masked off
                            len = akb3_scan_end(A.T, B.text, cs + i, cs + (lane >= 30 ? 62 : lane == 29 ? 94 : 96), kc, B.off, B.n_rows, r_lo, r_hi) - (cs + i);
                        if (len > 0x7FFFF) len = 0x7FFFF;
                        ev[k++] = pos | (kc << 11) | ((uint32_t)len << 13);
                    }
                }
            }
            __syncwarp();
            // ---- pass 1: tokens per event
            int running = 0;
            for (int base = 0; base < E; base += 32) {
                const int e = base + lane;
                int n = 0;
                uint32_t rr = 0;
                if (e < E) {
                    const uint32_t v = ev[e];
                    const int64_t p = ws + (v & 1023u);
                    if (v & (1u << 10)) {
                        rr = res[e];
                        n = (int)(rr & 0x3FFFFFFFu);
                    } else {
                        const uint32_t kc = (v >> 11) & 3u;
                        uint32_t len = v >> 13;
                        if (len == 0x7FFFFu) len = 0xFFFFFu;                 // clamped: the cold paths look for the end again
                        long long hit = -1, slot = -1;
                        unsigned long long h = 0, want = 0;
                        const bool cacheable = len <= AKW_MAXLEN && A.C.e != nullptr;
                        unsigned long long tag = 0;
                        if (cacheable) hit = akw_lookup(A.C, B.text, p, len, B.text_end, &h, &want, &slot, &tag);
                        if (hit >= 0) {
                            n = (int)((tag & AKW_NTOK_MASK) >> 3);
                            rr = ((uint32_t)hit << 5) | (uint32_t)n;
                        } else {
                            rr = akb_event_slow(A, B, p, len, kc, h, want, slot, cacheable, r_lo, r_hi, st);
                            n = (rr & 0x80000000u) ? (int)(rr & 0x3FFFFFFFu) : (int)(rr & 31u);
                        }
                    }
                }
                int inc = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                    if (lane >= d) inc += y;
                }
                if (e < E) {
                    res[e] = rr;
                    eoff[e] = (uint16_t)(running + inc - n);
                }
                running += __shfl_sync(0xFFFFFFFFu, inc, 31);
            }
            total = running;
            if (total >= 65536) use_list = false;
        }
        if (use_list) {
            __syncwarp();
            const int first_total = E480 < E ? (int)eoff[E480] : total;
            unsigned int toff = 0;
            if (lane == 0) toff = atomicAdd(&s_cursor, (unsigned int)total);
            toff = __shfl_sync(0xFFFFFFFFu, toff, 0);
            const bool fits = (int64_t)toff + total <= A.slice_cap;
            if (lane == 0) {
                A.wt_total[wt0] = first_total;
                A.wt_toff[wt0] = slice + toff;
                if (two) {
                    A.wt_total[wt0 + 1] = total - first_total;
                    A.wt_toff[wt0 + 1] = slice + toff + first_total;
                }
                if (!fits) st |= AK_ST_OVERFLOW;
            }
            // ---- pass 2: ids from the cache entries, row splits relative to their warp tile
            int32_t* tbase = A.temp + slice + toff;
            for (int base = 0; base < E; base += 32) {
                const int e = base + lane;
                if (e >= E) continue;
                const uint32_t v = ev[e], rr = res[e];
                const int o = eoff[e];
                const int64_t p = ws + (v & 1023u);
                if (v & (1u << 10)) {
                    const int rel = (v & 1023u) >= (uint32_t)AKF_WARP_BYTES ? first_total : 0;
                    int64_t g = r_w0 + (v >> 13);
                    int k = o;
                    while (g <= B.n_rows && B.off[g] == p) {
                        if (g > 0 && A.M.eos >= 0) { if (fits) tbase[k] = A.M.eos; ++k; }
                        A.id_splits[g] = k - rel;
                        if (g < B.n_rows && A.M.bos >= 0) { if (fits) tbase[k] = A.M.bos; ++k; }
                        ++g;
                    }
                } else if (fits) {
                    const int n = (int)(rr & 31u);
                    if (!(rr & 0x80000000u)) {
                        const unsigned long long* en = A.C.e + (unsigned long long)((rr >> 5) & 0x3FFFFu) * AKW_ENTRY;
#pragma unroll 1
                        for (int i = 0; i < n; i += 2) {
                            const unsigned long long q = akw_ldc(en + 8 + (i >> 1));
                            tbase[o + i] = (int32_t)(uint32_t)q;
                            if (i + 1 < n) tbase[o + i + 1] = (int32_t)(uint32_t)(q >> 32);
                        }
                    } else {
                        uint32_t len = v >> 13;
                        if (len == 0x7FFFFu) len = 0xFFFFFu;
                        akb_event_write_direct(A, B, p, len, (v >> 11) & 3u, r_lo, r_hi, tbase + o, (int64_t)(rr & 0x3FFFFFFFu));
                    }
                }
            }
            __syncwarp();
            ak_raise(B.result, st);
        } else {
            // too many events for the list: the two warp tiles one after the other through the v2 routine
            ak_raise(B.result, st & ~(uint32_t)AK_ST_OVERFLOW);
            __syncwarp();
            akb_v2_warp_tile(A, B, wt0, lane, lut, ev, res, eoff, &s_cursor, slice);
            if (two) akb_v2_warp_tile(A, B, wt0 + 1, lane, lut, ev, res, eoff, &s_cursor, slice);
        }
    }
}


// move every warp tile's block to its final place and make the row splits global
__global__ void __launch_bounds__(AKW_GROUP) ak_bf_copy_kernel(const AkBfArgs A) {
    __shared__ int ws[33];
    __shared__ int s_excl[AKW_GROUP / 32][33];
    __shared__ long long s_delta[AKW_GROUP / 32][32];
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    const int n_wt = akw_n_tiles(B, A.base0);
    const int n_groups = (n_wt + AKW_GROUP - 1) / AKW_GROUP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int gidx = blockIdx.x; gidx < n_groups; gidx += gridDim.x) {
        const int t = gidx * AKW_GROUP + tid;
        const int mine = t < n_wt ? A.wt_total[t] : 0;
        int total;
        const int pre = ak_block_exscan<AKW_GROUP>(mine, ws, total);
        const int64_t dst = A.sum_base[gidx] + pre;                   // final position of this lane's warp tile
        // the rows that start in my warp tile: splits made global
        if (t < n_wt) {
            const int64_t r0 = A.wrow[t], r1 = A.wrow[t + 1];
            for (int64_t r = r0; r < r1 && r <= B.n_rows; ++r) A.id_splits[r] += dst;
        }
        // the warp's 32 blocks are contiguous in the output: one flat, coalesced copy
        const int pre_w = __shfl_sync(0xFFFFFFFFu, pre, 0);
        s_excl[warp][lane] = pre - pre_w;
        s_delta[warp][lane] = (t < n_wt ? A.wt_toff[t] : 0) - (long long)(pre - pre_w);
        const int W = __shfl_sync(0xFFFFFFFFu, pre + mine, 31) - pre_w;
        if (lane == 0) s_excl[warp][32] = 0x7FFFFFFF;
        __syncwarp();
        const int64_t dst0 = __shfl_sync(0xFFFFFFFFu, dst, 0);
        if (dst0 + W > A.id_cap) { if (lane == 0 && W > 0) ak_raise(B.result, AK_ST_OVERFLOW); }
        else akw_flat_copy<int32_t>(A.temp, A.ids, dst0, W, s_excl[warp], s_delta[warp], lane);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// K4b Unigram encode  (reference tokenizer.py:191): one row per thread, Viterbi ring in registers / local memory,
// final back-pointers in a global scratch (4 B per code point), ids written backwards from the row's end.
// ------------------------------------------------------------------------------------------------
struct AkUniArgs {
    AkBatch B;
    AkUniDev U;
    uint32_t* back;            // scratch: row r uses back[(off[r] - text_begin) + 2 r ...]
    int32_t* ids;
    int64_t id_cap;
    int64_t* id_splits;
};

__global__ void __launch_bounds__(AK_ROWS_BLOCK) ak_unigram_kernel(const AkUniArgs A) {
    __shared__ int ws[33];
    __shared__ int s_tile;
    __shared__ long long s_base;
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    for (;;) {
        const int tile = ak_next_tile(B.ticket, &s_tile);
        if (tile >= B.n_tiles) break;
        const int64_t r = (int64_t)tile * AK_ROWS_BLOCK + threadIdx.x;
        int64_t n = 0, cnt = 0;
        uint32_t* back = nullptr;
        if (r < B.n_rows) {
            const int64_t rs = B.off[r], re = B.off[r + 1];
            back = A.back + (rs - B.text_begin) + 2 * r;
            n = ak_unigram_forward(A.U, B.text, rs, re, back);
            cnt = ak_unigram_backtrack(A.U, back, n, nullptr, 0, 0);
        }
        int total;
        const int pre = ak_block_exscan<AK_ROWS_BLOCK>((int)cnt, ws, total);
        if (threadIdx.x < 32) {
            long long b = ak_tile_prefix(B.state0, tile, total, (unsigned int*)&B.result[2], AK_ST_SPIN);
            if (threadIdx.x == 0) {
                s_base = b;
                if (tile == B.n_tiles - 1) B.totals[0] = b + total;
            }
        }
        __syncthreads();
        if (r < B.n_rows) {
            const int64_t obase = s_base + pre;
            A.id_splits[r] = obase;
            if (r == B.n_rows - 1) A.id_splits[B.n_rows] = obase + cnt;
            if (obase + cnt > A.id_cap) ak_raise(B.result, AK_ST_OVERFLOW);
            ak_unigram_backtrack(A.U, back, n, A.ids, obase + cnt, A.id_cap);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1b roman_phonetic_signature  (reference normalize.py:59-89): one word per row, one row per thread.
// lower() [all scripts, Final_Sigma] -> collapse runs >= 3 -> ee$ -> i, oo$ -> u -> aa kh gh ch th ph bh dh.
// ------------------------------------------------------------------------------------------------
struct AkSigArgs {
    AkBatch B;
    AkTables T;
    uint32_t* cps;             // scratch: one uint32 per input byte
    uint8_t* out;
    int64_t out_cap;
    int64_t* out_off;
};

__global__ void __launch_bounds__(AK_ROWS_BLOCK) ak_signature_kernel(const AkSigArgs A) {
    __shared__ int ws[33];
    __shared__ int s_tile;
    __shared__ long long s_base;
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    for (;;) {
        const int tile = ak_next_tile(B.ticket, &s_tile);
        if (tile >= B.n_tiles) break;
        const int64_t r = (int64_t)tile * AK_ROWS_BLOCK + threadIdx.x;
        int n = 0, cnt = 0;
        uint32_t* a = nullptr;
        if (r < B.n_rows) {
            const int64_t rs = B.off[r], re = B.off[r + 1];
            a = A.cps + (rs - B.text_begin);
            n = ak_signature_row(A.T, B.text, rs, re, a);
            for (int i = 0; i < n; ++i) cnt += ak_utf8_len(a[i]);
        }
        int total;
        const int pre = ak_block_exscan<AK_ROWS_BLOCK>(cnt, ws, total);
        if (threadIdx.x < 32) {
            long long b = ak_tile_prefix(B.state0, tile, total, (unsigned int*)&B.result[2], AK_ST_SPIN);
            if (threadIdx.x == 0) {
                s_base = b;
                if (tile == B.n_tiles - 1) B.totals[0] = b + total;
            }
        }
        __syncthreads();
        if (r < B.n_rows) {
            const int64_t obase = s_base + pre;
            A.out_off[r] = obase;
            if (r == B.n_rows - 1) A.out_off[B.n_rows] = obase + cnt;
            if (obase + cnt <= A.out_cap) {
                uint8_t* o = A.out + obase;
                for (int i = 0; i < n; ++i) o += ak_encode(a[i], o);
            } else {
                ak_raise(B.result, AK_ST_OVERFLOW);
            }
        }
    }
}

// ================================================================================================
// host side: context, model upload, C ABI
// ================================================================================================
struct akshar_ctx {
    int device = 0;
    int sm_count = 148;
    AkTables T{};
    std::vector<void*> allocs;
    std::string err;
    int64_t launches = 0;
    bool has_bpe = false, has_uni = false;
    AkBpeHost bpe_h;
    AkBpeDev bpe_d{};
    AkUniHost uni_h;
    AkUniDev uni_d{};
    std::vector<void*> bpe_allocs, uni_allocs;
    AkWordCache wc{};              // working copy, restored from wc_image at the start of every BPE call
    unsigned long long* wc_image = nullptr;
    size_t wc_bytes = 0;
    int occ_bf = 0, occ_sf = 0, occ_bf3 = 0;
    // optional CUDA-event timing of the dominant kernel of each stage (bench.py's roofline line)
    bool timing = false;
    bool wc_hold = false;          // akshar_word_cache_hold: skip the per-call restore of the word cache
    cudaEvent_t tev[AKSHAR_TIMER_COUNT][2] = {};
    bool tev_valid[AKSHAR_TIMER_COUNT] = {};
    int occ_norm = 0, occ_seg = 0, occ_bpe = 0, occ_uni = 0, occ_sig = 0, occ_nf_classify = 0, occ_nf_write = 0, occ_nf3 = 0, occ_sf3 = 0;
};

#define AK_CUDA(ctx, call)                                                                         \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
            return AKSHAR_E_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

template <class V>
static int ak_upload(akshar_ctx* ctx, std::vector<void*>& owner, const V* src, size_t n, const V** dst) {
    void* d = nullptr;
    size_t bytes = (n ? n : 1) * sizeof(V);
    AK_CUDA(ctx, cudaMalloc(&d, bytes));
    owner.push_back(d);
    if (n) AK_CUDA(ctx, cudaMemcpy(d, src, n * sizeof(V), cudaMemcpyHostToDevice));
    *dst = (const V*)d;
    return 0;
}

extern "C" {

int akshar_version(void) { return 100; }

const char* akshar_status_str(int code) {
    switch (code) {
        case AKSHAR_OK: return "ok";
        case AKSHAR_E_ARG: return "bad argument";
        case AKSHAR_E_CUDA: return "CUDA error";
        case AKSHAR_E_MODEL: return "model could not be parsed or uses an unsupported configuration";
        case AKSHAR_E_NOMODEL: return "no model loaded for this encoder";
        case AKSHAR_E_WORKSPACE: return "workspace too small";
        default: return "unknown";
    }
}

int akshar_ctx_create(int device, akshar_ctx** out) {
    if (!out) return AKSHAR_E_ARG;
    *out = nullptr;
    akshar_ctx* ctx = new (std::nothrow) akshar_ctx();
    if (!ctx) return AKSHAR_E_ARG;
    ctx->device = device;
    *out = ctx;      // returned even on failure so that akshar_last_error can be read; caller destroys it
    AK_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    AK_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        ctx->err = "akshar_b200 is built for sm_100a (B200); device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        return AKSHAR_E_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    int rc;
#define UP(field, arr, n, type) \
    if ((rc = ak_upload<type>(ctx, ctx->allocs, (const type*)(arr), (size_t)(n), (const type**)&ctx->T.field))) return rc;
    UP(page_index, ak_tbl_page_index, AK_N_PAGES, uint16_t)
    UP(leaves, ak_tbl_leaves, AK_N_LEAF_PAGES * 256, uint32_t)
    UP(decomp_keys, ak_tbl_decomp_keys, AK_N_DECOMP, uint32_t)
    UP(decomp_off, ak_tbl_decomp_off, AK_N_DECOMP + 1, uint16_t)
    UP(decomp_data, ak_tbl_decomp_data, AK_N_DECOMP_DATA, uint32_t)
    UP(pair_keys, ak_tbl_pair_keys, AK_N_PAIRS, unsigned long long)
    UP(pair_vals, ak_tbl_pair_vals, AK_N_PAIRS, uint32_t)
    UP(ll_keys, ak_tbl_latin_lower_keys, AK_N_LATIN_LOWER, uint32_t)
    UP(ll_vals, ak_tbl_latin_lower_vals, AK_N_LATIN_LOWER, uint32_t)
    UP(fl_keys, ak_tbl_full_lower_keys, AK_N_FULL_LOWER, uint32_t)
    UP(fl_vals, ak_tbl_full_lower_vals, AK_N_FULL_LOWER, uint32_t)
#undef UP
    ctx->T.n_decomp = AK_N_DECOMP;
    ctx->T.n_pairs = AK_N_PAIRS;
    ctx->T.n_ll = AK_N_LATIN_LOWER;
    ctx->T.n_fl = AK_N_FULL_LOWER;
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_norm, ak_normalize_kernel, AK_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_nf_classify, ak_nf_classify_kernel, AK_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_nf_write, ak_nf_write_kernel, AK_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_nf3, ak_nf3_classify_kernel, AKN3_THREADS, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_bf, ak_bf_encode_kernel, AK_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_bf3, ak_bf3_encode_kernel, AKB3_THREADS, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_sf, ak_sf_kernel, AK_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_sf3, ak_sf3_kernel, AKS3_THREADS, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_seg, ak_segment_kernel, AK_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_bpe, ak_bpe_kernel, AK_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_uni, ak_unigram_kernel, AK_ROWS_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_sig, ak_signature_kernel, AK_ROWS_BLOCK, 0));
    return AKSHAR_OK;
}

static void ak_free_list(std::vector<void*>& v) {
    for (void* p : v) cudaFree(p);
    v.clear();
}

void akshar_ctx_destroy(akshar_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ak_free_list(ctx->allocs);
    ak_free_list(ctx->bpe_allocs);
    ak_free_list(ctx->uni_allocs);
    for (int i = 0; i < AKSHAR_TIMER_COUNT; ++i)
        for (int k = 0; k < 2; ++k)
            if (ctx->tev[i][k]) cudaEventDestroy(ctx->tev[i][k]);
    delete ctx;
}

const char* akshar_last_error(akshar_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int akshar_word_cache_hold(akshar_ctx* ctx, int hold) {
    if (!ctx) return AKSHAR_E_ARG;
    ctx->wc_hold = hold != 0;
    return AKSHAR_OK;
}

int akshar_timing_enable(akshar_ctx* ctx, int enable) {
    if (!ctx) return AKSHAR_E_ARG;
    AK_CUDA(ctx, cudaSetDevice(ctx->device));
    if (enable)
        for (int i = 0; i < AKSHAR_TIMER_COUNT; ++i)
            for (int k = 0; k < 2; ++k)
                if (!ctx->tev[i][k]) AK_CUDA(ctx, cudaEventCreate(&ctx->tev[i][k]));
    ctx->timing = enable != 0;
    for (int i = 0; i < AKSHAR_TIMER_COUNT; ++i) ctx->tev_valid[i] = false;
    return AKSHAR_OK;
}

int akshar_timing_read(akshar_ctx* ctx, int timer, float* ms) {
    if (!ctx || !ms || timer < 0 || timer >= AKSHAR_TIMER_COUNT) return AKSHAR_E_ARG;
    if (!ctx->tev_valid[timer]) return AKSHAR_E_ARG;
    AK_CUDA(ctx, cudaEventSynchronize(ctx->tev[timer][1]));
    AK_CUDA(ctx, cudaEventElapsedTime(ms, ctx->tev[timer][0], ctx->tev[timer][1]));
    return AKSHAR_OK;
}

int64_t akshar_launch_count(akshar_ctx* ctx) { return ctx ? ctx->launches : 0; }

// workspace layout (all regions 256-byte aligned):
//   [0, 256)            control block: tickets[8] (int), changed flag, pool cursor
//   state regions       4 x n_tiles x 8 bytes (two counters x two passes)
//   scratch             max(unigram back-pointers 4 (n_bytes + 2 n_rows + 2), signature code points 4 n_bytes,
//                           BPE: NFC'd text n_bytes + n_bytes / 8 + 1024, its row offsets 8 (n_rows + 1), long-word pool)
static inline size_t ak_align(size_t x) { return (x + 255) & ~(size_t)255; }
static inline int64_t ak_tiles_of(int64_t n_bytes, int64_t n_rows) {
    int64_t a = (n_bytes + n_bytes / 8 + 1024 + 32 + AKF_TILE - 1) / AKF_TILE;     // fast-kernel tiles; covers the NFC'd copy too
    int64_t b = (n_rows + AK_ROWS_BLOCK - 1) / AK_ROWS_BLOCK;
    return (a > b ? a : b) + 1;
}
static inline size_t ak_pool_ints(int64_t n_bytes) {
    int64_t p = n_bytes / 2 + (1 << 16);
    return (size_t)p;
}
struct AkWsLayout {
    size_t control, state, tile_row, nfc_text, nfc_off, pool, bf_tiles, bf_temp, scratch, total;
    int64_t bf_temp_cap;
    int64_t nfc_cap;
};
static AkWsLayout ak_ws_layout(int64_t n_bytes, int64_t n_rows) {
    AkWsLayout L;
    size_t tiles = (size_t)ak_tiles_of(n_bytes, n_rows);
    L.control = 0;
    L.state = 256;
    L.tile_row = L.state + ak_align(4 * tiles * 8);
    size_t at = L.tile_row + ak_align(8 * (tiles * AKF_WARPS + 2));
    L.scratch = at;
    size_t uni = 4 * (size_t)(n_bytes + 2 * n_rows + 2);
    L.nfc_cap = n_bytes + n_bytes / 8 + 1024;
    L.nfc_text = at;
    L.nfc_off = L.nfc_text + ak_align((size_t)L.nfc_cap);
    L.pool = L.nfc_off + ak_align(8 * (size_t)(n_rows + 1));
    L.bf_tiles = L.pool + ak_align(4 * ak_pool_ints(n_bytes));
    {
        const size_t nwt = tiles * AKF_WARPS + 8, ng = nwt / AKW_GROUP + 2;
        L.bf_temp = L.bf_tiles + ak_align((nwt + 2) * 8) + ak_align(nwt * 4) + ak_align(nwt * 8) + ak_align(ng * 4) + ak_align((ng + 1) * 8);
    }
    L.bf_temp_cap = n_bytes / 2 + 2 * n_rows + 1024;
    size_t bpe = (L.bf_temp + ak_align(4 * (size_t)L.bf_temp_cap)) - at;
    // fast normalize: per-lane info words, tile totals / bases, slow work list
    size_t nf = ak_align(tiles * AK_BLOCK * 4) + ak_align(tiles * 4) + ak_align((tiles + 1) * 8) +
                ak_align((tiles * AK_BLOCK / 16 + 1024) * sizeof(AkSlowEntry));
    // fast segment: tile totals / offsets / bases for two streams + the temporary streams
    const size_t sf_nwt = tiles * AKF_WARPS + 8, sf_ng = sf_nwt / AKW_GROUP + 2;
    size_t sf = ak_align((sf_nwt + 2) * 8) + 2 * ak_align(sf_nwt * 4) + 2 * ak_align(sf_nwt * 8) + 2 * ak_align(sf_ng * 4) +
                2 * ak_align((sf_ng + 1) * 8) +
                ak_align((size_t)(n_bytes / 2 + n_rows + 1024) * 4) + ak_align((size_t)(n_bytes / 8 + n_rows + 1024) * 5);
    size_t m = uni > bpe ? uni : bpe;
    if (sf > m) m = sf;
    L.total = at + ak_align(m > nf ? m : nf);
    return L;
}

size_t akshar_workspace_bytes(int64_t n_bytes, int64_t n_rows) {
    if (n_bytes < 0 || n_rows < 0) return 0;
    return ak_ws_layout(n_bytes, n_rows).total;
}

struct AkCall {
    akshar_ctx* ctx;
    AkBatch B;
    AkWsLayout L;
    char* ws;
    size_t ws_bytes;       // what the caller really passed: the temporary streams use everything beyond the fixed part
    cudaStream_t stream;
};

// validates the common arguments, clears the control block + result, fills AkBatch (tiles for `mode`)
static int ak_begin(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows, int64_t text_begin,
                    int64_t text_end, int mode, int rows_block, int64_t* d_result, void* d_workspace, size_t workspace_bytes,
                    void* stream, AkCall& C) {
    if (!ctx) return AKSHAR_E_ARG;
    if (!d_row_offsets || !d_result || n_rows < 0 || text_end < text_begin || (!d_text && text_end > text_begin) ||
        (mode != AKSHAR_MODE_TILES && mode != AKSHAR_MODE_ROWS)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    const int64_t n_bytes = text_end - text_begin;
    C.ctx = ctx;
    C.L = ak_ws_layout(n_bytes, n_rows);
    if (!d_workspace || workspace_bytes < C.L.total) {
        ctx->err = "workspace too small: need " + std::to_string(C.L.total) + " bytes";
        return AKSHAR_E_WORKSPACE;
    }
    C.ws = (char*)d_workspace;
    C.ws_bytes = workspace_bytes;
    C.stream = (cudaStream_t)stream;
    AK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t tiles = (size_t)ak_tiles_of(n_bytes, n_rows);
    AK_CUDA(ctx, cudaMemsetAsync(C.ws, 0, 256 + ak_align(4 * tiles * 8), C.stream));
    AK_CUDA(ctx, cudaMemsetAsync(d_result, 0, 4 * sizeof(int64_t), C.stream));
    AkBatch& B = C.B;
    B.text = d_text;
    B.off = d_row_offsets;
    B.n_rows = n_rows;
    B.text_begin = text_begin;
    B.text_end = text_end;
    B.mode = mode;
    if (rows_block > 0) B.n_tiles = (int)((n_rows + rows_block - 1) / rows_block);
    else if (mode == AKSHAR_MODE_TILES) B.n_tiles = (int)((n_bytes + 1 + AK_TILE - 1) / AK_TILE);
    else B.n_tiles = (int)((n_rows + AK_BLOCK - 1) / AK_BLOCK);
    B.ticket = (int*)C.ws;
    B.state0 = (unsigned long long*)(C.ws + C.L.state);
    B.state1 = B.state0 + tiles;
    B.result = d_result;
    B.totals = d_result;
    B.run_if = nullptr;
    B.dyn_end = nullptr;
    return AKSHAR_OK;
}

// brackets one kernel launch with events when timing is enabled
struct AkTimed {
    akshar_ctx* ctx;
    int slot;
    cudaStream_t s;
    AkTimed(akshar_ctx* c, int sl, cudaStream_t st) : ctx(c), slot(sl), s(st) {
        if (ctx->timing && ctx->tev[slot][0]) cudaEventRecord(ctx->tev[slot][0], s);
    }
    ~AkTimed() {
        if (ctx->timing && ctx->tev[slot][1]) { cudaEventRecord(ctx->tev[slot][1], s); ctx->tev_valid[slot] = true; }
    }
};

static int ak_grid(akshar_ctx* ctx, int occ, int n_tiles) {
    int g = ctx->sm_count * (occ > 0 ? occ : 1);
    return n_tiles < g ? (n_tiles > 0 ? n_tiles : 1) : g;
}

static int ak_after_launch(akshar_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->err = std::string(what) + " launch: " + cudaGetErrorString(e);
        return AKSHAR_E_CUDA;
    }
    ctx->launches++;
    return AKSHAR_OK;
}

// a batch with no rows: every ragged output is just splits[0] = 0
static int ak_empty_rows(akshar_ctx* ctx, int64_t* a, int64_t* b, cudaStream_t s) {
    if (a) AK_CUDA(ctx, cudaMemsetAsync(a, 0, sizeof(int64_t), s));
    if (b) AK_CUDA(ctx, cudaMemsetAsync(b, 0, sizeof(int64_t), s));
    return AKSHAR_OK;
}

// normalize_text launch: the fast kernel for the default flags in tile mode, the generic walker kernel otherwise
static int ak_run_normalize(akshar_ctx* ctx, AkCall& C, const AkBatch& B, uint32_t flags, uint8_t* out, int64_t out_cap,
                            int64_t* out_off) {
    const bool raw_fast = flags == AK_NORM_ROMAN;            // clean_hinglish=False: bit-stream kernel only
    if ((flags == (AK_NORM_ROMAN | AK_NORM_CLEAN) || raw_fast) && B.mode == AKSHAR_MODE_TILES && !B.dyn_end) {
        AkFastNormArgs F;
        F.flags = flags;
        F.B = B;
        F.T = ctx->T;
        F.out = out;
        F.out_cap = out_cap;
        F.out_off = out_off;
        F.base0 = B.text_begin - (int64_t)(((uintptr_t)B.text + (uintptr_t)B.text_begin) & 15u);
        F.B.n_tiles = (int)((B.text_end - F.base0 + AKF_TILE) / AKF_TILE);
        int64_t* tile_row = (int64_t*)(C.ws + C.L.tile_row);
        F.tile_row = tile_row;
        const int entries = F.B.n_tiles * AKF_WARPS + 1;       // one entry per warp tile (480 bytes)
        ak_warp_rows_kernel<<<(entries + 255) / 256, 256, 0, C.stream>>>(B, F.base0, entries, tile_row);
        int rc = ak_after_launch(ctx, "warp-rows");
        if (rc) return rc;
        AkNfWork W;
        const size_t nt = (size_t)F.B.n_tiles;
        char* wp = C.ws + C.L.scratch;
        W.info = (uint32_t*)wp;                         wp += ak_align(nt * AK_BLOCK * 4);
        W.tile_total = (int32_t*)wp;                    wp += ak_align(nt * 4);
        W.tile_base = (int64_t*)wp;                     wp += ak_align((nt + 1) * 8);
        W.slow = (AkSlowEntry*)wp;
        W.n_slow = (unsigned int*)(C.ws + 72);
        W.slow_cap = (unsigned int)(nt * AK_BLOCK / 16 + 1024);
        {
            AkTimed tm(ctx, AKSHAR_TIMER_NORMALIZE_CLASSIFY, C.stream);
            if (getenv("AKSHAR_NORM_V2") && !raw_fast)
                ak_nf_classify_kernel<<<ak_grid(ctx, ctx->occ_nf_classify, F.B.n_tiles), AK_BLOCK, 0, C.stream>>>(F, W);
            else
                ak_nf3_classify_kernel<<<ak_grid(ctx, ctx->occ_nf3, F.B.n_tiles), AKN3_THREADS, 0, C.stream>>>(F, W);
        }
        if ((rc = ak_after_launch(ctx, "normalize-classify"))) return rc;
        AkNfSlowArgs S;
        S.B = B;
        S.T = ctx->T;
        S.W = W;
        S.tile_row = tile_row;
        S.out = out;
        S.out_cap = out_cap;
        S.out_off = out_off;
        S.write = 0;
        S.flags = flags;
        const int slow_grid = ctx->sm_count * AKN_SLOW_MINB;      // latency bound: as many walkers in flight as fit
        ak_nf_slow_kernel<<<slow_grid, 128, 0, C.stream>>>(S);
        if ((rc = ak_after_launch(ctx, "normalize-slow-count"))) return rc;
        ak_nf_scan_kernel<<<1, 1024, 0, C.stream>>>(W.tile_total, W.tile_base, F.B.n_tiles, B.totals, B, F.base0);
        if ((rc = ak_after_launch(ctx, "normalize-scan"))) return rc;
        {
            AkTimed tm(ctx, AKSHAR_TIMER_NORMALIZE_WRITE, C.stream);
            ak_nf_write_kernel<<<ak_grid(ctx, ctx->occ_nf_write, F.B.n_tiles), AK_BLOCK, 0, C.stream>>>(F, W);
        }
        if ((rc = ak_after_launch(ctx, "normalize-write"))) return rc;
        S.write = 1;
        ak_nf_slow_kernel<<<slow_grid, 128, 0, C.stream>>>(S);
        return ak_after_launch(ctx, "normalize-slow-write");
    }
    AkNormArgs A;
    A.B = B;
    A.T = ctx->T;
    A.flags = flags;
    A.out = out;
    A.out_cap = out_cap;
    A.out_off = out_off;
    ak_normalize_kernel<<<ak_grid(ctx, ctx->occ_norm, A.B.n_tiles), AK_BLOCK, 0, C.stream>>>(A);
    return ak_after_launch(ctx, "normalize");
}

int akshar_normalize_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                           int64_t text_begin, int64_t text_end, uint32_t flags, int mode, uint8_t* d_out_text,
                           int64_t out_capacity, int64_t* d_out_row_offsets, int64_t* d_result, void* d_workspace,
                           size_t workspace_bytes, void* stream) {
    AkCall C;
    int rc = ak_begin(ctx, d_text, d_row_offsets, n_rows, text_begin, text_end, mode, 0, d_result, d_workspace, workspace_bytes,
                      stream, C);
    if (rc) return rc;
    if (!d_out_row_offsets || out_capacity < 0 || (!d_out_text && out_capacity > 0) || (flags & ~15u)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (n_rows == 0) return ak_empty_rows(ctx, d_out_row_offsets, nullptr, C.stream);
    return ak_run_normalize(ctx, C, C.B, flags, d_out_text, out_capacity, d_out_row_offsets);
}

int akshar_segment_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                         int64_t text_begin, int64_t text_end, uint32_t flags, int mode, int32_t* d_cluster_ends,
                         int64_t cluster_capacity, int64_t* d_cluster_splits, int32_t* d_run_ends, uint8_t* d_run_tags,
                         int64_t run_capacity, int64_t* d_run_splits, int64_t* d_result, void* d_workspace,
                         size_t workspace_bytes, void* stream) {
    AkCall C;
    int rc = ak_begin(ctx, d_text, d_row_offsets, n_rows, text_begin, text_end, mode, 0, d_result, d_workspace, workspace_bytes,
                      stream, C);
    if (rc) return rc;
    const bool want_c = (flags & AKSHAR_SEG_CLUSTERS) != 0, want_r = (flags & AKSHAR_SEG_RUNS) != 0;
    if ((flags & ~7u) || (!want_c && !want_r) || ((flags & AKSHAR_SEG_MATRAS) && !want_c) ||
        (want_c && (!d_cluster_splits || cluster_capacity < 0 || (!d_cluster_ends && cluster_capacity > 0))) ||
        (want_r && (!d_run_splits || run_capacity < 0 || ((!d_run_ends || !d_run_tags) && run_capacity > 0)))) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (n_rows == 0) return ak_empty_rows(ctx, want_c ? d_cluster_splits : nullptr, want_r ? d_run_splits : nullptr, C.stream);
    AkSegArgs A;
    A.B = C.B;
    A.T = ctx->T;
    A.flags = flags;
    A.o.cluster_ends = d_cluster_ends;
    A.o.cluster_splits = d_cluster_splits;
    A.o.run_ends = d_run_ends;
    A.o.run_tags = d_run_tags;
    A.o.run_splits = d_run_splits;
    A.o.cbase = A.o.rbase = 0;
    A.o.ccap = want_c ? cluster_capacity : 0;
    A.o.rcap = want_r ? run_capacity : 0;
    if (mode == AKSHAR_MODE_TILES) {
        const int64_t n_bytes = text_end - text_begin;
        const size_t tiles = (size_t)ak_tiles_of(n_bytes, n_rows);
        AkSfArgs F;
        F.B = C.B;
        F.T = ctx->T;
        F.flags = flags;
        F.base0 = text_begin - (int64_t)(((uintptr_t)d_text + (uintptr_t)text_begin) & 15u);
        const int nwt = (int)((text_end - F.base0 + AKF_WARP_BYTES) / AKF_WARP_BYTES);
        const int ngroups = (nwt + AKW_GROUP - 1) / AKW_GROUP;
        char* wp = C.ws + C.L.scratch;
        F.wrow = (int64_t*)wp;          wp += ak_align(((size_t)nwt + 2) * 8);
        F.c_total = (int32_t*)wp;       wp += ak_align((size_t)nwt * 4);
        F.r_total = (int32_t*)wp;       wp += ak_align((size_t)nwt * 4);
        F.c_toff = (int64_t*)wp;        wp += ak_align((size_t)nwt * 8);
        F.r_toff = (int64_t*)wp;        wp += ak_align((size_t)nwt * 8);
        F.c_sums = (int32_t*)wp;        wp += ak_align((size_t)ngroups * 4);
        F.r_sums = (int32_t*)wp;        wp += ak_align((size_t)ngroups * 4);
        F.c_sum_base = (int64_t*)wp;    wp += ak_align(((size_t)ngroups + 1) * 8);
        F.r_sum_base = (int64_t*)wp;    wp += ak_align(((size_t)ngroups + 1) * 8);
        const bool v2 = getenv("AKSHAR_SEG_V2") != nullptr;
        const int grid = v2 ? ak_grid(ctx, ctx->occ_sf, (nwt + AKF_WARPS - 1) / AKF_WARPS)
                            : ak_grid(ctx, ctx->occ_sf3, ((nwt + 1) / 2 + AKS3_THREADS / 32 - 1) / (AKS3_THREADS / 32));
        int64_t tc_cap, tr_cap;
        {
            // the temporary streams share what is left of the workspace: 4 B per cluster end, 5 B per run end
            const size_t left = C.ws_bytes - (size_t)(wp - C.ws) - 1024;
            if (want_c && want_r) { tc_cap = (int64_t)(left * 3 / 4 / 4); tr_cap = (int64_t)(left / 4 / 5); }
            else if (want_c) { tc_cap = (int64_t)(left / 4); tr_cap = 0; }
            else { tc_cap = 0; tr_cap = (int64_t)(left / 5); }
        }
        F.c_slice = tc_cap / grid;
        F.r_slice = tr_cap / grid;
        F.tc = (int32_t*)wp;            wp += ak_align((size_t)tc_cap * 4);
        F.tr = (int32_t*)wp;            wp += ak_align((size_t)tr_cap * 4);
        F.tt = (uint8_t*)wp;
        F.o = A.o;
        ak_warp_rows_kernel<<<(nwt + 2 + 255) / 256, 256, 0, C.stream>>>(C.B, F.base0, nwt + 2, (int64_t*)F.wrow);
        if ((rc = ak_after_launch(ctx, "segment-warp-rows"))) return rc;
        {
            AkTimed tm(ctx, AKSHAR_TIMER_SEGMENT, C.stream);
            if (v2) ak_sf_kernel<<<grid, AK_BLOCK, 0, C.stream>>>(F);
            else ak_sf3_kernel<<<grid, AKS3_THREADS, 0, C.stream>>>(F);
        }
        if ((rc = ak_after_launch(ctx, "segment-fast"))) return rc;
        if (want_c) {
            ak_wt_sums_kernel<<<ak_grid(ctx, 8, ngroups), AKW_GROUP, 0, C.stream>>>(C.B, F.base0, F.c_total, F.c_sums);
            if ((rc = ak_after_launch(ctx, "segment-sums"))) return rc;
            ak_nf_scan_kernel<<<1, 1024, 0, C.stream>>>(F.c_sums, F.c_sum_base, ngroups, d_result, C.B, F.base0, AKF_WARP_BYTES * AKW_GROUP);
            if ((rc = ak_after_launch(ctx, "segment-scan"))) return rc;
        }
        if (want_r) {
            ak_wt_sums_kernel<<<ak_grid(ctx, 8, ngroups), AKW_GROUP, 0, C.stream>>>(C.B, F.base0, F.r_total, F.r_sums);
            if ((rc = ak_after_launch(ctx, "segment-sums"))) return rc;
            ak_nf_scan_kernel<<<1, 1024, 0, C.stream>>>(F.r_sums, F.r_sum_base, ngroups, d_result + 1, C.B, F.base0, AKF_WARP_BYTES * AKW_GROUP);
            if ((rc = ak_after_launch(ctx, "segment-scan"))) return rc;
        }
        ak_sf_copy_kernel<<<ak_grid(ctx, 8, ngroups), AKW_GROUP, 0, C.stream>>>(F);
        return ak_after_launch(ctx, "segment-copy");
    }
    ak_segment_kernel<<<ak_grid(ctx, ctx->occ_seg, A.B.n_tiles), AK_BLOCK, 0, C.stream>>>(A);
    return ak_after_launch(ctx, "segment");
}

int akshar_signature_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                           int64_t text_begin, int64_t text_end, uint8_t* d_out_text, int64_t out_capacity,
                           int64_t* d_out_row_offsets, int64_t* d_result, void* d_workspace, size_t workspace_bytes,
                           void* stream) {
    AkCall C;
    int rc = ak_begin(ctx, d_text, d_row_offsets, n_rows, text_begin, text_end, AKSHAR_MODE_ROWS, AK_ROWS_BLOCK, d_result,
                      d_workspace, workspace_bytes, stream, C);
    if (rc) return rc;
    if (!d_out_row_offsets || out_capacity < 0 || (!d_out_text && out_capacity > 0)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (n_rows == 0) return ak_empty_rows(ctx, d_out_row_offsets, nullptr, C.stream);
    AkSigArgs A;
    A.B = C.B;
    A.T = ctx->T;
    A.cps = (uint32_t*)(C.ws + C.L.scratch);
    A.out = d_out_text;
    A.out_cap = out_capacity;
    A.out_off = d_out_row_offsets;
    ak_signature_kernel<<<ak_grid(ctx, ctx->occ_sig, A.B.n_tiles), AK_ROWS_BLOCK, 0, C.stream>>>(A);
    return ak_after_launch(ctx, "signature");
}

int akshar_load_bpe_json(akshar_ctx* ctx, const char* json, size_t len) {
    if (!ctx || !json) return AKSHAR_E_ARG;
    AkBpeHost h;
    std::string e = ak_parse_bpe_json(json, len, h);
    if (!e.empty()) {
        ctx->err = e;
        return AKSHAR_E_MODEL;
    }
    AK_CUDA(ctx, cudaSetDevice(ctx->device));
    AK_CUDA(ctx, cudaDeviceSynchronize());
    ak_free_list(ctx->bpe_allocs);
    ctx->has_bpe = false;
    AkBpeDev d{};
    int rc;
    if ((rc = ak_upload<int32_t>(ctx, ctx->bpe_allocs, h.cp_direct.data(), h.cp_direct.size(), &d.cp_direct))) return rc;
    if ((rc = ak_upload<uint32_t>(ctx, ctx->bpe_allocs, h.cp_keys.data(), h.cp_keys.size(), &d.cp_keys))) return rc;
    if ((rc = ak_upload<int32_t>(ctx, ctx->bpe_allocs, h.cp_ids.data(), h.cp_ids.size(), &d.cp_ids))) return rc;
    if ((rc = ak_upload<unsigned long long>(ctx, ctx->bpe_allocs, h.mkeys.data(), h.mkeys.size(), &d.mkeys))) return rc;
    if ((rc = ak_upload<unsigned long long>(ctx, ctx->bpe_allocs, h.mvals.data(), h.mvals.size(), &d.mvals))) return rc;
    d.n_cp = (int)h.cp_keys.size();
    d.mbits = h.mbits;
    d.bos = h.bos;
    d.eos = h.eos;
    // word cache image: every vocabulary string that is exactly one pre-tokenizer word, encoded by the merge loop
    {
        AkBpeDev hm{};
        hm.cp_direct = h.cp_direct.data();
        hm.cp_keys = h.cp_keys.data();
        hm.cp_ids = h.cp_ids.data();
        hm.n_cp = (int)h.cp_keys.size();
        hm.mkeys = h.mkeys.data();
        hm.mvals = h.mvals.data();
        hm.mbits = h.mbits;
        hm.bos = h.bos;
        hm.eos = h.eos;
        AkTables ht{};
        ht.page_index = ak_tbl_page_index;
        ht.leaves = ak_tbl_leaves;
        const uint32_t bits = 18;
        std::vector<unsigned long long> img((size_t)AKW_ENTRY << bits, 0ull);
        AkWordCache hc;
        hc.e = img.data();
        hc.bits = bits;
        std::vector<int32_t> poolbuf(4096);
        unsigned long long used = 0;
        AkPool hp;
        hp.base = poolbuf.data();
        hp.used = &used;
        hp.cap = poolbuf.size();
        for (size_t id = 0; id < h.id_to_token.size(); ++id) {
            const std::string& tok = h.id_to_token[id];
            if (tok.empty() || tok.size() > AKW_MAXLEN || h.is_special[id]) continue;
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
            int32_t out_ids[AKW_MAXTOK + 1];
            sink.buf = out_ids;
            sink.cap = AKW_MAXTOK + 1;
            sink.stride = 1;
            sink.cnt = 0;
            sink.direct = false;
            sink.gout = nullptr;
            sink.gbase = 0;
            sink.gcap = 0;
            ak_bpe_word(hm, ht, tb, 0, n, k, sink, hp, st);
            if (st || sink.cnt > AKW_MAXTOK) continue;
            const unsigned long long hh = akw_hash(tb, 0, (uint32_t)n);
            const unsigned long long want = akw_want(hh, (uint32_t)n);
            long long slot;
            if (akw_find(hc, hh, want, tb, 0, (uint32_t)n, &slot) < 0 && slot >= 0)
                akw_insert(hc, slot, want, tb, 0, (uint32_t)n, out_ids, sink.cnt);
        }
        const unsigned long long* dimg = nullptr;
        if ((rc = ak_upload<unsigned long long>(ctx, ctx->bpe_allocs, img.data(), img.size(), &dimg))) return rc;
        void* work = nullptr;
        AK_CUDA(ctx, cudaMalloc(&work, img.size() * 8));
        ctx->bpe_allocs.push_back(work);
        ctx->wc_image = (unsigned long long*)dimg;
        ctx->wc.e = (unsigned long long*)work;
        ctx->wc.bits = bits;
        ctx->wc_bytes = img.size() * 8;
    }
    ctx->bpe_d = d;
    ctx->bpe_h = std::move(h);
    ctx->has_bpe = true;
    return AKSHAR_OK;
}

int akshar_load_spm_model(akshar_ctx* ctx, const void* proto, size_t len) {
    if (!ctx || !proto) return AKSHAR_E_ARG;
    AkUniHost h;
    std::string e = ak_parse_spm_model(proto, len, h);
    if (!e.empty()) {
        ctx->err = e;
        return AKSHAR_E_MODEL;
    }
    AK_CUDA(ctx, cudaSetDevice(ctx->device));
    AK_CUDA(ctx, cudaDeviceSynchronize());
    ak_free_list(ctx->uni_allocs);
    ctx->has_uni = false;
    AkUniDev d{};
    int rc;
    if ((rc = ak_upload<unsigned long long>(ctx, ctx->uni_allocs, h.tkeys.data(), h.tkeys.size(), &d.tkeys))) return rc;
    if ((rc = ak_upload<unsigned long long>(ctx, ctx->uni_allocs, h.tvals.data(), h.tvals.size(), &d.tvals))) return rc;
    {
        std::vector<unsigned long long> kv(2 * h.tkeys.size());
        for (size_t i = 0; i < h.tkeys.size(); ++i) { kv[2 * i] = h.tkeys[i]; kv[2 * i + 1] = h.tvals[i]; }
        if ((rc = ak_upload<unsigned long long>(ctx, ctx->uni_allocs, kv.data(), kv.size(), &d.tkv))) return rc;
    }
    if ((rc = ak_upload<float>(ctx, ctx->uni_allocs, h.score.data(), h.score.size(), &d.score))) return rc;
    if ((rc = ak_upload<uint8_t>(ctx, ctx->uni_allocs, h.usable.data(), h.usable.size(), &d.usable))) return rc;
    if ((rc = ak_upload<int32_t>(ctx, ctx->uni_allocs, h.byte_id, 256, &d.byte_id))) return rc;
    d.tbits = h.tbits;
    d.unk_id = h.unk_id;
    d.unk_score = h.unk_score;
    d.flags = h.flags;
    ctx->uni_d = d;
    ctx->uni_h = std::move(h);
    ctx->has_uni = true;
    return AKSHAR_OK;
}

int akshar_vocab_size(akshar_ctx* ctx, int kind) {
    if (!ctx) return AKSHAR_E_ARG;
    if (kind == 0) return ctx->has_bpe ? ctx->bpe_h.vocab_size : AKSHAR_E_NOMODEL;
    if (kind == 1) return ctx->has_uni ? (int)ctx->uni_h.piece.size() : AKSHAR_E_NOMODEL;
    return AKSHAR_E_ARG;
}

int akshar_vocab_token(akshar_ctx* ctx, int kind, int id, const char** bytes, int* len, int* type) {
    if (!ctx || !bytes || !len || !type) return AKSHAR_E_ARG;
    if (kind == 0) {
        if (!ctx->has_bpe) return AKSHAR_E_NOMODEL;
        if (id < 0 || (size_t)id >= ctx->bpe_h.id_to_token.size()) return AKSHAR_E_ARG;
        *bytes = ctx->bpe_h.id_to_token[(size_t)id].data();
        *len = (int)ctx->bpe_h.id_to_token[(size_t)id].size();
        *type = ctx->bpe_h.is_special[(size_t)id];
        return AKSHAR_OK;
    }
    if (kind == 1) {
        if (!ctx->has_uni) return AKSHAR_E_NOMODEL;
        if (id < 0 || (size_t)id >= ctx->uni_h.piece.size()) return AKSHAR_E_ARG;
        *bytes = ctx->uni_h.piece[(size_t)id].data();
        *len = (int)ctx->uni_h.piece[(size_t)id].size();
        *type = ctx->uni_h.type[(size_t)id];
        return AKSHAR_OK;
    }
    return AKSHAR_E_ARG;
}

// the three BPE launches over batch B (B may carry dyn_end / run_if from an earlier stage of a pipeline)
static int ak_run_bpe(akshar_ctx* ctx, AkCall& C, const AkBatch& B, int64_t max_bytes, int32_t* d_ids, int64_t id_capacity,
                      int64_t* d_id_splits) {
    int rc;
    const size_t tiles = (size_t)ak_tiles_of(max_bytes, B.n_rows);
    int* tickets = (int*)C.ws;
    unsigned int* changed = (unsigned int*)(C.ws + 64);
    AkPool pool;
    pool.base = (int32_t*)(C.ws + C.L.pool);
    pool.used = (unsigned long long*)(C.ws + 128);
    pool.cap = ak_pool_ints(max_bytes);
    // A workspace larger than the minimum gives half of the surplus to the long-word pool (the other half enlarges the
    // temporary id stream): a batch full of words beyond AK_BPE_LOCAL symbols raises AKSHAR_ST_WORD with the minimum,
    // the caller grows the workspace and calls again.  The pool then lives at the workspace's tail.
    size_t pool_tail = 0;
    if (C.ws_bytes > C.L.total + (1u << 20)) {
        pool_tail = ((C.ws_bytes - C.L.total) / 2) & ~(size_t)255;
        if (pool_tail / 4 > pool.cap) {
            pool.base = (int32_t*)(C.ws + ((C.ws_bytes - pool_tail) & ~(size_t)255));
            pool.cap = pool_tail / 4 - 64;
        } else pool_tail = 0;
    }
    // pass 1: encode the text as it is; raises `changed` when some NFC segment is not already normalized
    AkBpeArgs A;
    A.B = B;
    A.B.ticket = tickets + 1;
    A.B.state0 = C.B.state0 + 1 * tiles;
    A.T = ctx->T;
    A.M = ctx->bpe_d;
    A.pool = pool;
    A.ids = d_ids;
    A.id_cap = id_capacity;
    A.id_splits = d_id_splits;
    A.changed = changed;
    const int bpe_tiles = B.dyn_end ? (int)tiles : B.n_tiles;
    if (B.mode == AKSHAR_MODE_TILES) {
        // fast kernels: word cache, no ordered tile dependency
        AkBfArgs F;
        F.B = B;
        F.T = ctx->T;
        F.M = ctx->bpe_d;
        F.C = ctx->wc;
        F.pool = pool;
        F.base0 = B.text_begin - (int64_t)(((uintptr_t)B.text + (uintptr_t)B.text_begin) & 15u);
        const int64_t span = (B.dyn_end ? max_bytes : B.text_end) - F.base0;
        const int nwt_ub = (int)((span + AKF_WARP_BYTES) / AKF_WARP_BYTES);
        const int ngroups_ub = (nwt_ub + AKW_GROUP - 1) / AKW_GROUP;
        char* wp = C.ws + C.L.bf_tiles;
        F.wrow = (int64_t*)wp;                       wp += ak_align(((size_t)nwt_ub + 2) * 8);
        F.wt_total = (int32_t*)wp;                   wp += ak_align((size_t)nwt_ub * 4);
        F.wt_toff = (int64_t*)wp;                    wp += ak_align((size_t)nwt_ub * 8);
        F.sums = (int32_t*)wp;                       wp += ak_align((size_t)ngroups_ub * 4);
        F.sum_base = (int64_t*)wp;                   wp += ak_align(((size_t)ngroups_ub + 1) * 8);
        F.temp = (int32_t*)wp;
        const bool v2 = getenv("AKSHAR_BPE_V2") != nullptr;
        const int grid = v2 ? ak_grid(ctx, ctx->occ_bf, (nwt_ub + AKF_WARPS - 1) / AKF_WARPS)
                            : ak_grid(ctx, ctx->occ_bf3, ((nwt_ub + 1) / 2 + AKB3_WARPS - 1) / AKB3_WARPS);
        F.slice_cap = (int64_t)((C.ws_bytes - pool_tail - 512 - (size_t)(wp - C.ws)) / 4 / (size_t)grid);      // a larger workspace = larger slices
        F.ids = d_ids;
        F.id_cap = id_capacity;
        F.id_splits = d_id_splits;
        F.changed = changed;
        if (!ctx->wc_hold)
            AK_CUDA(ctx, cudaMemcpyAsync(ctx->wc.e, ctx->wc_image, ctx->wc_bytes, cudaMemcpyDeviceToDevice, C.stream));
        ak_warp_rows_kernel<<<(nwt_ub + 2 + 255) / 256, 256, 0, C.stream>>>(B, F.base0, nwt_ub + 2, (int64_t*)F.wrow);
        if ((rc = ak_after_launch(ctx, "bpe-warp-rows"))) return rc;
        {
            AkTimed tm(ctx, AKSHAR_TIMER_BPE_ENCODE, C.stream);
            if (v2) ak_bf_encode_kernel<<<grid, AK_BLOCK, 0, C.stream>>>(F);
            else ak_bf3_encode_kernel<<<grid, AKB3_THREADS, 0, C.stream>>>(F);
        }
        if ((rc = ak_after_launch(ctx, "bpe-fast"))) return rc;
        ak_wt_sums_kernel<<<ak_grid(ctx, 8, ngroups_ub), AKW_GROUP, 0, C.stream>>>(B, F.base0, F.wt_total, F.sums);
        if ((rc = ak_after_launch(ctx, "bpe-sums"))) return rc;
        ak_nf_scan_kernel<<<1, 1024, 0, C.stream>>>(F.sums, F.sum_base, ngroups_ub, B.totals, B, F.base0, AKF_WARP_BYTES * AKW_GROUP);
        if ((rc = ak_after_launch(ctx, "bpe-scan"))) return rc;
        ak_bf_copy_kernel<<<ak_grid(ctx, 8, ngroups_ub), AKW_GROUP, 0, C.stream>>>(F);
        if ((rc = ak_after_launch(ctx, "bpe-copy"))) return rc;
    } else {
        ak_bpe_kernel<<<ak_grid(ctx, ctx->occ_bpe, bpe_tiles), AK_BLOCK, 0, C.stream>>>(A);
        if ((rc = ak_after_launch(ctx, "bpe"))) return rc;
    }
    // passes 2 + 3 (device-side conditional: both exit at once while `changed` is clear): NFC into the workspace,
    // then encode that copy over the same outputs.  HF's NFKC == NFC on the closed alphabet (the exotic spaces it
    // folds to U+0020 are all \\s and never reach a word).
    int64_t* nfc_total = (int64_t*)(C.ws + 192);
    AkNormArgs N;
    N.B = B;
    N.B.ticket = tickets + 2;
    N.B.state0 = C.B.state0 + 2 * tiles;
    N.B.totals = nfc_total;
    N.B.run_if = changed;
    N.T = ctx->T;
    N.flags = 0;
    N.out = (uint8_t*)(C.ws + C.L.nfc_text);
    N.out_cap = C.L.nfc_cap;
    N.out_off = (int64_t*)(C.ws + C.L.nfc_off);
    ak_normalize_kernel<<<ak_grid(ctx, ctx->occ_norm, bpe_tiles), AK_BLOCK, 0, C.stream>>>(N);
    if ((rc = ak_after_launch(ctx, "bpe-nfc"))) return rc;
    AkBpeArgs A2 = A;
    A2.B.text = N.out;
    A2.B.off = N.out_off;
    A2.B.text_begin = 0;
    A2.B.text_end = 0;
    A2.B.dyn_end = nfc_total;
    A2.B.run_if = changed;
    A2.B.ticket = tickets + 3;
    A2.B.state0 = C.B.state0 + 3 * tiles;
    A2.changed = (unsigned int*)(C.ws + 68);      // scratch flag: the copy is in NFC by construction
    ak_bpe_kernel<<<ak_grid(ctx, ctx->occ_bpe, (int)tiles), AK_BLOCK, 0, C.stream>>>(A2);
    return ak_after_launch(ctx, "bpe-renormalized");
}

static int ak_run_unigram(akshar_ctx* ctx, AkCall& C, const AkBatch& B, int64_t max_bytes, int32_t* d_ids,
                          int64_t id_capacity, int64_t* d_id_splits) {
    const size_t tiles = (size_t)ak_tiles_of(max_bytes, B.n_rows);
    AkUniArgs A;
    A.B = B;
    A.B.mode = AKSHAR_MODE_ROWS;
    A.B.n_tiles = (int)((B.n_rows + AK_ROWS_BLOCK - 1) / AK_ROWS_BLOCK);
    A.B.ticket = (int*)C.ws + 1;
    A.B.state0 = C.B.state0 + 1 * tiles;
    A.U = ctx->uni_d;
    A.back = (uint32_t*)(C.ws + C.L.scratch);
    A.ids = d_ids;
    A.id_cap = id_capacity;
    A.id_splits = d_id_splits;
    {
        AkTimed tm(ctx, AKSHAR_TIMER_UNIGRAM, C.stream);
        ak_unigram_kernel<<<ak_grid(ctx, ctx->occ_uni, A.B.n_tiles), AK_ROWS_BLOCK, 0, C.stream>>>(A);
    }
    return ak_after_launch(ctx, "unigram");
}

int akshar_encode_bpe_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                            int64_t text_begin, int64_t text_end, int mode, int32_t* d_ids, int64_t id_capacity,
                            int64_t* d_id_splits, int64_t* d_result, void* d_workspace, size_t workspace_bytes,
                            void* stream) {
    AkCall C;
    int rc = ak_begin(ctx, d_text, d_row_offsets, n_rows, text_begin, text_end, mode, 0, d_result, d_workspace, workspace_bytes,
                      stream, C);
    if (rc) return rc;
    if (!ctx->has_bpe) {
        ctx->err = "no BPE model loaded";
        return AKSHAR_E_NOMODEL;
    }
    if (!d_id_splits || id_capacity < 0 || (!d_ids && id_capacity > 0)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (n_rows == 0) return ak_empty_rows(ctx, d_id_splits, nullptr, C.stream);
    return ak_run_bpe(ctx, C, C.B, text_end - text_begin, d_ids, id_capacity, d_id_splits);
}

int akshar_encode_unigram_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                                int64_t text_begin, int64_t text_end, int mode, int32_t* d_ids, int64_t id_capacity,
                                int64_t* d_id_splits, int64_t* d_result, void* d_workspace, size_t workspace_bytes,
                                void* stream) {
    AkCall C;
    (void)mode;      // the lattice is always built row by row
    int rc = ak_begin(ctx, d_text, d_row_offsets, n_rows, text_begin, text_end, AKSHAR_MODE_ROWS, AK_ROWS_BLOCK, d_result,
                      d_workspace, workspace_bytes, stream, C);
    if (rc) return rc;
    if (!ctx->has_uni) {
        ctx->err = "no Unigram model loaded";
        return AKSHAR_E_NOMODEL;
    }
    if (!d_id_splits || id_capacity < 0 || (!d_ids && id_capacity > 0)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (n_rows == 0) return ak_empty_rows(ctx, d_id_splits, nullptr, C.stream);
    return ak_run_unigram(ctx, C, C.B, text_end - text_begin, d_ids, id_capacity, d_id_splits);
}

int akshar_tokenizer_encode_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                                  int64_t text_begin, int64_t text_end, uint32_t norm_flags, int kind, int mode,
                                  uint8_t* d_norm_text, int64_t norm_capacity, int64_t* d_norm_row_offsets, int32_t* d_ids,
                                  int64_t id_capacity, int64_t* d_id_splits, int64_t* d_result, void* d_workspace,
                                  size_t workspace_bytes, void* stream) {
    if (!ctx) return AKSHAR_E_ARG;
    if (norm_capacity < 0 || !d_norm_row_offsets || (!d_norm_text && norm_capacity > 0) || (norm_flags & ~15u) ||
        (kind != 0 && kind != 1) || !d_id_splits || id_capacity < 0 || (!d_ids && id_capacity > 0)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (kind == 0 ? !ctx->has_bpe : !ctx->has_uni) {
        ctx->err = "no model loaded for this encoder";
        return AKSHAR_E_NOMODEL;
    }
    const int64_t n_bytes = text_end - text_begin;
    const int64_t max_bytes = n_bytes > norm_capacity ? n_bytes : norm_capacity;
    // the workspace must cover both stages: validate against the larger text
    if (workspace_bytes < ak_ws_layout(max_bytes, n_rows).total) {
        ctx->err = "workspace too small: need " + std::to_string(ak_ws_layout(max_bytes, n_rows).total) + " bytes";
        return AKSHAR_E_WORKSPACE;
    }
    AkCall C;
    int rc = ak_begin(ctx, d_text, d_row_offsets, n_rows, text_begin, text_end, mode, 0, d_result, d_workspace, workspace_bytes,
                      stream, C);
    if (rc) return rc;
    C.L = ak_ws_layout(max_bytes, n_rows);
    const size_t tiles = (size_t)ak_tiles_of(max_bytes, n_rows);
    AK_CUDA(ctx, cudaMemsetAsync(C.ws, 0, 256 + ak_align(4 * tiles * 8), C.stream));
    C.B.state0 = (unsigned long long*)(C.ws + C.L.state);
    C.B.state1 = C.B.state0 + tiles;
    if (n_rows == 0) return ak_empty_rows(ctx, d_norm_row_offsets, d_id_splits, C.stream);
    // stage 1: normalize_text (tokenizer.py:185 preprocess); its byte total lands in result[1]
    AkBatch B1 = C.B;
    B1.totals = d_result + 1;
    if ((rc = ak_run_normalize(ctx, C, B1, norm_flags, d_norm_text, norm_capacity, d_norm_row_offsets))) return rc;
    // stage 2: the model on the normalized rows; their length is only known on the device (dyn_end)
    AkBatch B2 = C.B;
    B2.text = d_norm_text;
    B2.off = d_norm_row_offsets;
    B2.text_begin = 0;
    B2.text_end = 0;
    B2.dyn_end = d_result + 1;
    if (mode != AKSHAR_MODE_TILES) B2.n_tiles = (int)((n_rows + AK_BLOCK - 1) / AK_BLOCK);
    if (kind == 0) return ak_run_bpe(ctx, C, B2, max_bytes, d_ids, id_capacity, d_id_splits);
    return ak_run_unigram(ctx, C, B2, max_bytes, d_ids, id_capacity, d_id_splits);
}

}  // extern "C"
