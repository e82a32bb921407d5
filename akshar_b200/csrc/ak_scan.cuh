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
