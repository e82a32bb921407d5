// Per-sentence statistics and the cluster-merging feature wrappers, computed from the segment kernel's outputs without
// leaving the device:
//   ak_comp_kernel      analyze_text_composition (reference segment.py:210-236): one warp per row -> five counts
//   ak_cm_flag_kernel / ak_cm_keep_kernel / ak_cm_write_kernel
//                       akshara_level_tokenization (features.py:28-55) and preserve_nukta (features.py:173-206): both are
//                       "drop some boundaries of segment_akshars(text)" -- which ones depends only on whether a cluster
//                       holds U+094D / U+093C and, for the nukta rule, on its place in a run of such clusters
#pragma once

#define AKF_STATS 5           // per row: akshars, script runs, code points, code points in devanagari runs, in roman runs

struct AkCompArgs {
    const uint8_t* text;
    const int64_t* off;
    int64_t n_rows;
    const int64_t* cluster_splits;
    const int32_t* run_ends;          // row-relative byte offsets
    const uint8_t* run_tags;
    const int64_t* run_splits;
    int32_t* stats;                   // [n_rows * AKF_STATS]
};

__global__ void __launch_bounds__(256) ak_comp_kernel(const AkCompArgs A) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < A.n_rows; r += warps) {
        const int64_t rs = A.off[r], re = A.off[r + 1];
        const int64_t k0 = A.run_splits[r], k1 = A.run_splits[r + 1];
        int total = 0, dev = 0, rom = 0;
        int64_t k = k0;                                       // run of the first byte of this warp step (uniform)
        for (int64_t p0 = rs; p0 < re; p0 += 32) {
            const int64_t p = p0 + lane;
            const bool lead = p < re && (A.text[p] & 0xC0u) != 0x80u;
            // runs are few per row: advance from the step's first run
            int64_t kk = k;
            const int32_t rel = (int32_t)(p - rs);
            while (kk < k1 - 1 && A.run_ends[kk] <= rel) ++kk;
            const uint32_t tag = (lead && kk < k1) ? A.run_tags[kk] : 255u;
            total += __popc(__ballot_sync(0xFFFFFFFFu, lead));
            dev += __popc(__ballot_sync(0xFFFFFFFFu, lead && tag == 0u));
            rom += __popc(__ballot_sync(0xFFFFFFFFu, lead && tag == 1u));
            const int32_t rel_next = (int32_t)(p0 + 32 - rs);
            while (k < k1 - 1 && A.run_ends[k] <= rel_next) ++k;
        }
        if (lane == 0) {
            int32_t* o = A.stats + r * AKF_STATS;
            o[0] = (int32_t)(A.cluster_splits[r + 1] - A.cluster_splits[r]);
            o[1] = (int32_t)(k1 - k0);
            o[2] = total;
            o[3] = dev;
            o[4] = rom;
        }
    }
}

// ---- cluster merging -------------------------------------------------------------------------------------------------
#define AKCM_AKSHARA 0        // features.py:28-55: consecutive clusters that hold a halant (U+094D) are one akshara
#define AKCM_NUKTA 1          // features.py:173-206: a cluster that holds a nukta (U+093C) takes the next cluster with it

struct AkCmArgs {
    const uint8_t* text;
    const int64_t* off;
    int64_t n_rows;
    const int32_t* ends;              // cluster END byte offsets, row-relative
    const int64_t* splits;            // [n_rows + 1]
    int64_t n;                        // clusters
    int rule;
    uint8_t* flag;                    // [n]: bit 0 holds the code point, bit 1 last cluster of its row
    int32_t* count;                   // [tiles]
    const int64_t* base;              // [tiles]
    int32_t* out_ends;
    int64_t* out_splits;
    int64_t cap;
    int64_t* result;
};
#define AKCM_THREADS 256
#define AKCM_PER 4
#define AKCM_TILE (AKCM_THREADS * AKCM_PER)

__global__ void __launch_bounds__(256) ak_cm_flag_kernel(const AkCmArgs A) {
    const uint8_t b2 = A.rule == AKCM_AKSHARA ? 0xA5u : 0xA4u, b3 = A.rule == AKCM_AKSHARA ? 0x8Du : 0xBCu;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += (int64_t)gridDim.x * blockDim.x) {
        // row of cluster i: last r with splits[r] <= i
        int64_t lo = 0, hi = A.n_rows;
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (A.splits[mid] <= i) lo = mid; else hi = mid;
        }
        const int64_t r = lo;
        const int64_t rs = A.off[r];
        const int64_t s = rs + (i == A.splits[r] ? 0 : A.ends[i - 1]), e = rs + A.ends[i];
        uint8_t f = (i + 1 == A.splits[r + 1]) ? 2u : 0u;
        for (int64_t q = s; q + 2 < e; ++q)
            if (A.text[q] == 0xE0u && A.text[q + 1] == b2 && A.text[q + 2] == b3) { f |= 1u; break; }
        A.flag[i] = f;
    }
}

// is the boundary after cluster i kept?
__device__ __forceinline__ bool akcm_keep(const AkCmArgs& A, int64_t i) {
    const uint8_t f = A.flag[i];
    if (!(f & 1u) || (f & 2u)) return true;
    if (A.rule == AKCM_AKSHARA) return !(A.flag[i + 1] & 1u);
    // nukta: clusters pair up from the start of every run of nukta clusters; the first of a pair loses its boundary
    int64_t j = i;
    while (j > 0 && (A.flag[j - 1] & 3u) == 1u) --j;
    return ((i - j) & 1) != 0;
}

template <bool WRITE>
__global__ void __launch_bounds__(AKCM_THREADS) ak_cm_kernel(const AkCmArgs A) {
    __shared__ int ws[33];
    const int64_t n_tiles = (A.n + AKCM_TILE - 1) / AKCM_TILE;
    uint32_t st = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * AKCM_TILE + (int64_t)threadIdx.x * AKCM_PER;
        bool keep[AKCM_PER];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < AKCM_PER; ++k) {
            keep[k] = i0 + k < A.n && akcm_keep(A, i0 + k);
            sum += keep[k] ? 1 : 0;
        }
        int total;
        const int pre = ak_block_exscan<AKCM_THREADS>(sum, ws, total);
        if (!WRITE) {
            if (threadIdx.x == 0) A.count[tile] = total;
            continue;
        }
        int64_t at = A.base[tile] + pre;
#pragma unroll
        for (int k = 0; k < AKCM_PER; ++k) {
            const int64_t i = i0 + k;
            if (i >= A.n) break;
            // a row's last cluster is always kept: the merged clusters of rows before r are the kept ones before splits[r]
            if (keep[k]) {
                if (at < A.cap) A.out_ends[at] = A.ends[i];
                else st |= AK_ST_OVERFLOW;
                ++at;
                if (A.flag[i] & 2u) {
                    // rows that end here: the next row(s) start at `at`
                    int64_t lo = 0, hi = A.n_rows;
                    while (hi - lo > 1) {
                        const int64_t mid = (lo + hi) >> 1;
                        if (A.splits[mid] <= i) lo = mid; else hi = mid;
                    }
                    for (int64_t r = lo + 1; r <= A.n_rows && A.splits[r] == i + 1; ++r) A.out_splits[r] = at;
                }
            }
        }
    }
    if (WRITE && blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t r = 0; r <= A.n_rows && A.splits[r] == 0; ++r) A.out_splits[r] = 0;      // empty rows in front
    }
    ak_raise(A.result, st);
}
