// Kernels of the on-device decode (cores and rationale: ak_decode.cuh).
//   ak_dec_mark_kernel     one thread per row: the position-dependent marks of its ids
//   ak_dec_kernel<count>   output bytes of every tile of AKD_TILE ids
//   (ak_scan_counts_kernel) tile bases, total
//   ak_dec_kernel<write>   lengths again, CTA scan, every id's bytes at their final place
//   ak_dec_rowoff_kernel   the row offsets of the text
// Algorithmic bytes: 2 x 4 per id read (ids, twice) + 1 mark written and read twice + the text written once.
#pragma once
#include "ak_decode.cuh"

#define AKD_THREADS 256
#define AKD_PER 4
#define AKD_TILE (AKD_THREADS * AKD_PER)

template <class IdT>
struct AkDecArgs {
    AkDecTable D;
    int form;
    const IdT* ids;
    int64_t n_ids;
    const int64_t* splits;             // [n_rows + 1] positions in ids
    int64_t n_rows;
    uint8_t* mark;                     // [n_ids], zeroed
    int32_t* count;                    // [n_tiles]
    const int64_t* base;               // [n_tiles]
    int32_t* tpre;                     // [ceil(n_ids / AKD_PER)] output bytes of the tile before each group of AKD_PER ids
    uint8_t* out;
    int64_t cap;
    int64_t* out_off;                  // [n_rows + 1]
    int64_t* result;
};

template <class IdT>
__global__ void __launch_bounds__(256) ak_dec_mark_kernel(const AkDecArgs<IdT> A) {
    uint32_t st = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < A.n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = A.splits[r], hi = A.splits[r + 1];
        if (lo < 0 || hi > A.n_ids || lo > hi) { st |= AK_ST_INTERNAL; continue; }
        akd_mark_row(A.D, A.form, A.ids, lo, hi, A.mark, st);
    }
    ak_raise(A.result, st);
}

template <class IdT, bool WRITE>
__global__ void __launch_bounds__(AKD_THREADS) ak_dec_kernel(const AkDecArgs<IdT> A) {
    __shared__ int ws[33];
    const int64_t n_tiles = (A.n_ids + AKD_TILE - 1) / AKD_TILE;
    uint32_t st = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i0 = tile * AKD_TILE + (int64_t)threadIdx.x * AKD_PER;
        uint32_t len[AKD_PER];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < AKD_PER; ++k) {
            len[k] = i0 + k < A.n_ids ? akd_emit(A.D, A.ids, A.mark, A.n_ids, i0 + k, nullptr, st) : 0u;
            sum += (int)len[k];
        }
        int total;
        const int pre = ak_block_exscan<AKD_THREADS>(sum, ws, total);
        if (!WRITE) {
            if (threadIdx.x == 0) A.count[tile] = total;
            continue;
        }
        int64_t at = A.base[tile] + pre;
        if (i0 < A.n_ids) A.tpre[i0 / AKD_PER] = pre;
#pragma unroll
        for (int k = 0; k < AKD_PER; ++k) {
            const int64_t i = i0 + k;
            if (i >= A.n_ids) break;
            if (len[k]) {
                if (at + len[k] <= A.cap) akd_emit(A.D, A.ids, A.mark, A.n_ids, i, A.out + at, st);
                else st |= AK_ST_OVERFLOW;
            }
            at += len[k];
        }
    }
    ak_raise(A.result, st);
}

// out_off[r] = output position of the first id of row r: tile base + its group's prefix + the lengths of the ids before it in the group
template <class IdT>
__global__ void __launch_bounds__(256) ak_dec_rowoff_kernel(const AkDecArgs<IdT> A) {
    uint32_t st = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= A.n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = A.splits[r];
        if (p >= A.n_ids) { A.out_off[r] = A.result[0]; continue; }
        const int64_t tile = p / AKD_TILE;
        int64_t at = A.base[tile];
        at += A.tpre[p / AKD_PER];
        for (int64_t i = p - p % AKD_PER; i < p; ++i) at += akd_emit(A.D, A.ids, A.mark, A.n_ids, i, nullptr, st);
        A.out_off[r] = at;
    }
}
