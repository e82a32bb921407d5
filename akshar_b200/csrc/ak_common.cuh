// Common kernel plumbing: the batch descriptor, span / tile bookkeeping, row-start masks of a lane, guarded edge loads.
#pragma once
#define AK_BLOCK 256
#define AK_SPAN 32
#define AK_TILE (AK_BLOCK * AK_SPAN)
#define AK_LOOKBACK_LIMIT 4096          // bytes a span may walk backwards in AKSHAR_MODE_TILES
#define AK_ROWS_BLOCK 128               // rows per tile for the row-per-thread kernels
#define AKF_WARPS (AK_BLOCK / 32)
#define AKF_TILE (AKF_WARPS * AKF_WARP_BYTES)     // 3840 text bytes per CTA tile in the fast kernels

static_assert((int)AK_ST_OVERFLOW == (int)AKSHAR_ST_OVERFLOW && (int)AK_ST_NFC_SEGMENT == (int)AKSHAR_ST_NFC_SEGMENT &&
              (int)AK_ST_PATHOLOGICAL == (int)AKSHAR_ST_PATHOLOGICAL && (int)AK_ST_ALPHABET == (int)AKSHAR_ST_ALPHABET &&
              (int)AK_ST_SPIN == (int)AKSHAR_ST_SPIN && (int)AK_ST_WORD == (int)AKSHAR_ST_WORD &&
              (int)AK_ST_INTERNAL == (int)AKSHAR_ST_INTERNAL && (int)AK_ST_BAD_ID == (int)AKSHAR_ST_BAD_ID, "status bits out of sync");
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
// Warp tiles.  The bit-stream kernels are warp-autonomous: a warp owns 960 text bytes (30 real lanes of 32 bytes + 2 halo
// lanes), finds the rows that start in them with shuffles and counts / writes its outputs without a CTA barrier or a
// global atomic, so a slow lane only delays its own warp.  Where outputs are compacted, a count pass and a scan over the
// per-warp-tile totals give every warp its final position before the emit pass runs.
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

// Slow chunks are not processed where they are found: a lane that cannot take the fast lane appends its chunk to a
// work list, and two small kernels run the exact walker over that list with one thread per entry.  A 16-byte walk
// costs tens of microseconds of dependent instructions; inside the tile kernels it would stall its whole CTA (and,
// through an ordered tile prefix, every later tile), on the list thousands of them overlap.

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

// the same, plus what the event-stream kernels need to number their rows without touching the offsets again: the index of
// the first row that starts in the lane's 32 bytes and how many rows start there (more than the mask's bits when rows are empty)
__device__ __forceinline__ uint32_t akn3_lane_rows2(const int64_t* off, int64_t n_rows, int64_t r_w0, int64_t ws, int lane,
                                                    int64_t& first_row, int& nrows) {
    uint32_t rows = 0;
    int64_t first = 0x7FFFFFFFFFFFFFFFll;
    int cntl = 0;
    const int64_t lo = ws - 32, hi = ws + AKN3_WARP_BYTES + 32;
    for (int64_t r = r_w0;; r += 32) {
        const int64_t mr = r + lane;
        const int64_t p = mr <= n_rows ? off[mr] : hi;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, p < hi);
        const int cnt = __popc(m);
        for (int j = 0; j < cnt; ++j) {
            const int rel = (int)(__shfl_sync(0xFFFFFFFFu, p, j) - lo);
            if ((rel >> 5) == lane) {
                rows |= 1u << (rel & 31);
                ++cntl;
                if (r + j < first) first = r + j;
            }
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
            if ((rel >> 5) == lane) {
                rows |= 1u << (rel & 31);
                ++cntl;
                if (r - j < first) first = r - j;
            }
        }
        if (cnt < 32) break;
    }
    first_row = first;
    nrows = cntl;
    return rows;
}

// the lane's 32 bytes at the edge of the text: bytes [lo, hi) of [cs, cs + 32), zero elsewhere.  Fully unrolled so that
// x[] stays in registers (a pointer handed to a non-inlined helper would put the lane's bytes of EVERY warp in local
// memory); only the first and the last warp tile of a text ever come here
__device__ __forceinline__ void akn3_load_edge(const uint8_t* text, int64_t cs, int lo, int hi, uint32_t* x) {
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = 4 * w + b;
            if (i >= lo && i < hi) v |= (uint32_t)text[cs + i] << (8 * b);
        }
        x[w] = v;
    }
}
