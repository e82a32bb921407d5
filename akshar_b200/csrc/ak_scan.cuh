// Block scan + single-pass ordered tile prefix (decoupled look-back) shared by every kernel.
//
// All batch kernels are PERSISTENT: gridDim = SMs x resident CTAs, each CTA draws tile numbers from an atomic
// ticket so that tile k only ever waits on tiles that are already running or finished (no dependence on the
// hardware's CTA dispatch order).  A tile publishes its output count as soon as it has counted, then sums the
// published counts of its predecessors 32 at a time until it meets one that already carries an inclusive
// prefix.  One 64-bit word per tile and counter: [63:62] flag, [61:0] value -- written with one store, so no
// fence is needed between value and flag.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define AK_FLAG_AGG (1ull << 62)
#define AK_FLAG_PREFIX (2ull << 62)
#define AK_FLAG_MASK (3ull << 62)
#define AK_SPIN_LIMIT (1 << 24)

__device__ __forceinline__ unsigned long long ak_ld_state(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ak_st_state(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// exclusive scan of one int per thread over the CTA; `ws` has 33 ints; total returned to every thread
template <int BLOCK>
__device__ __forceinline__ int ak_block_exscan(int v, int* ws, int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = lane < BLOCK / 32 ? ws[lane] : 0;
        int xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xFFFFFFFFu, xi, d);
            if (lane >= d) xi += y;
        }
        ws[lane] = xi - x;
        if (lane == 31) ws[32] = xi;
    }
    __syncthreads();
    int res = inc - v + ws[w];
    total = ws[32];
    __syncthreads();
    return res;
}

// the same over one 64-bit value per thread (several packed counters scanned at once); `ws` has 33 entries
template <int BLOCK>
__device__ __forceinline__ unsigned long long ak_block_exscan64(unsigned long long v, unsigned long long* ws, unsigned long long& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        const unsigned long long x = lane < BLOCK / 32 ? ws[lane] : 0ull;
        unsigned long long xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, xi, d);
            if (lane >= d) xi += y;
        }
        ws[lane] = xi - x;
        if (lane == 31) ws[32] = xi;
    }
    __syncthreads();
    const unsigned long long res = inc - v + ws[w];
    total = ws[32];
    __syncthreads();
    return res;
}

// Called by ALL 32 lanes of warp 0 with the same arguments.  Publishes `aggregate` for `tile` and returns the sum
// of the aggregates of tiles [0, tile).
__device__ __forceinline__ long long ak_tile_prefix(unsigned long long* state, int tile, long long aggregate,
                                                   unsigned int* status_word, unsigned int spin_bit) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) ak_st_state(state, AK_FLAG_PREFIX | (unsigned long long)aggregate);
        return 0;
    }
    if (lane == 0) ak_st_state(state + tile, AK_FLAG_AGG | (unsigned long long)aggregate);
    long long excl = 0;
    int look = tile - 1;
    for (;;) {
        const int idx = look - lane;
        unsigned long long s = AK_FLAG_PREFIX;      // virtual tiles before tile 0: prefix 0
        int spins = 0;
        if (idx >= 0) {
            s = ak_ld_state(state + idx);
            while ((s & AK_FLAG_MASK) == 0) {
                __nanosleep(40);
                s = ak_ld_state(state + idx);
                if (++spins > AK_SPIN_LIMIT) {       // never expected; do not hang the GPU
                    atomicOr(status_word, spin_bit);
                    s = AK_FLAG_PREFIX;
                }
            }
        }
        const unsigned pm = __ballot_sync(0xFFFFFFFFu, (s & AK_FLAG_MASK) == AK_FLAG_PREFIX);
        const int first = __ffs(pm) - 1;             // nearest predecessor that already has an inclusive prefix
        long long v = (pm == 0 || lane <= first) ? (long long)(s & ~AK_FLAG_MASK) : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        excl += v;
        if (pm) break;
        look -= 32;
    }
    if (lane == 0) ak_st_state(state + tile, AK_FLAG_PREFIX | (unsigned long long)(excl + aggregate));
    return excl;
}

// ------------------------------------------------------------------------------------------------
// Scans over the per-warp-tile aggregates of the event-stream encoders (one entry per 960 text bytes: ~1.1 M per GiB).
// Single pass, decoupled look-back over tiles of AKS_TILE entries -- a few hundred tiles, so the look-back chain is short
// and every tile does the same work.
// ------------------------------------------------------------------------------------------------
#define AKS_THREADS 256
#define AKS_PER 16
#define AKS_TILE (AKS_THREADS * AKS_PER)

// exclusive prefix sums of int32 counts -> int64 bases; *total receives the sum
__global__ void __launch_bounds__(AKS_THREADS) ak_scan_counts_kernel(const int32_t* v, long long n_fixed, const long long* n_dyn, long long n_mul,
                                                                      int64_t* base, int64_t* total, int* ticket, unsigned long long* state,
                                                                      unsigned int* status_word) {
    __shared__ int ws[33];
    __shared__ int s_tile;
    __shared__ long long s_base;
    const long long n = n_dyn ? *n_dyn * n_mul : n_fixed;
    const int n_tiles = (int)((n + AKS_TILE - 1) / AKS_TILE);
    const int tid = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(ticket, 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= n_tiles) break;
        const long long i0 = (long long)tile * AKS_TILE + (long long)tid * AKS_PER;
        int x[AKS_PER];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < AKS_PER; ++k) {
            x[k] = i0 + k < n ? v[i0 + k] : 0;
            sum += x[k];
        }
        int tot;
        const int pre = ak_block_exscan<AKS_THREADS>(sum, ws, tot);
        if (tid < 32) {
            const long long b = ak_tile_prefix(state, tile, tot, status_word, 16u);
            if (tid == 0) {
                s_base = b;
                if (tile == n_tiles - 1) *total = b + tot;
            }
        }
        __syncthreads();
        long long at = s_base + pre;
#pragma unroll
        for (int k = 0; k < AKS_PER; ++k) {
            if (i0 + k < n) base[i0 + k] = at;
            at += x[k];
        }
    }
    if (n == 0 && blockIdx.x == 0 && tid == 0) *total = 0;
}

// segmented sums: entry = (flag << 32) | float bits, (a (+) b) = (a.f | b.f, b.f ? b.v : a.v + b.v); out[i] = the float
// value of the exclusive prefix at i (the sum since the last flagged entry before i)
#define AKS_SEG_FLAG (1ull << 32)
__device__ __forceinline__ unsigned long long aks_seg_op(unsigned long long a, unsigned long long b) {
    if (b & AKS_SEG_FLAG) return b;
    return (a & AKS_SEG_FLAG) | (unsigned long long)__float_as_uint(__uint_as_float((uint32_t)a) + __uint_as_float((uint32_t)b));
}
__global__ void __launch_bounds__(AKS_THREADS) ak_scan_seg_kernel(const unsigned long long* v, const long long* n_dyn, float* out, int* ticket,
                                                                   unsigned long long* state, unsigned int* status_word) {
    __shared__ unsigned long long s_w[AKS_THREADS / 32];
    __shared__ int s_tile;
    __shared__ float s_base;
    const long long n = *n_dyn;
    const int n_tiles = (int)((n + AKS_TILE - 1) / AKS_TILE);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(ticket, 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= n_tiles) break;
        const long long i0 = (long long)tile * AKS_TILE + (long long)tid * AKS_PER;
        unsigned long long x[AKS_PER];
        unsigned long long agg = 0ull;                      // identity: no flag, 0.0f
#pragma unroll
        for (int k = 0; k < AKS_PER; ++k) {
            x[k] = i0 + k < n ? v[i0 + k] : 0ull;
            agg = aks_seg_op(agg, x[k]);
        }
        // warp inclusive scan of the thread aggregates
        unsigned long long inc = agg;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc = aks_seg_op(y, inc);
        }
        unsigned long long excl = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
        if (lane == 0) excl = 0ull;
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        unsigned long long wex = 0ull;                      // exclusive over the warps before mine
        for (int w = 0; w < warp; ++w) wex = aks_seg_op(wex, s_w[w]);
        if (tid == 0) {
            unsigned long long tagg = 0ull;
            for (int w = 0; w < AKS_THREADS / 32; ++w) tagg = aks_seg_op(tagg, s_w[w]);
            // look-back: state = [63:62] 1 aggregate / 2 inclusive prefix, [32] flag, [31:0] float
            float ex = 0.f;
            const unsigned long long mine = (tagg & AKS_SEG_FLAG) | (uint32_t)tagg;
            if (tile == 0) {
                ak_st_state(state, (2ull << 62) | mine);
            } else {
                ak_st_state(state + tile, (1ull << 62) | mine);
                bool done = false;
                for (int q = tile - 1; q >= 0 && !done; --q) {
                    unsigned long long sv = ak_ld_state(state + q);
                    int spins = 0;
                    while ((sv >> 62) == 0ull) {
                        __nanosleep(40);
                        sv = ak_ld_state(state + q);
                        if (++spins > AK_SPIN_LIMIT) { atomicOr(status_word, 16u); sv = (2ull << 62) | AKS_SEG_FLAG; }
                    }
                    ex += __uint_as_float((uint32_t)sv);
                    if ((sv & AKS_SEG_FLAG) || (sv >> 62) == 2ull) done = true;
                }
                const float incl = (tagg & AKS_SEG_FLAG) ? __uint_as_float((uint32_t)tagg) : ex + __uint_as_float((uint32_t)tagg);
                ak_st_state(state + tile, (2ull << 62) | (tagg & AKS_SEG_FLAG) | (unsigned long long)__float_as_uint(incl));
            }
            s_base = ex;
        }
        __syncthreads();
        // exclusive prefix of my first entry: tile base (+) warps before (+) threads before
        unsigned long long p = (unsigned long long)__float_as_uint(s_base);
        p = aks_seg_op(p, wex);
        p = aks_seg_op(p, excl);
#pragma unroll
        for (int k = 0; k < AKS_PER; ++k) {
            if (i0 + k < n) out[i0 + k] = __uint_as_float((uint32_t)p);
            p = aks_seg_op(p, x[k]);
        }
    }
}
