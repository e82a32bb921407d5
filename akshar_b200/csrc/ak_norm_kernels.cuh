// K1 normalize_text kernels (reference normalize.py:117-148): the generic span walker kernel and the bit-stream fast path
// (classify -> slow-lane walkers -> scan -> write).
#pragma once
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


#define AK_SLOW_BYTES 72
struct AkSlowEntry {
    int64_t pos;         // span start (absolute byte index)
    int64_t out_base;    // filled by the write kernel: where this span's output starts
    int32_t cnt;         // filled by the slow kernel's first pass
    int32_t tile;        // the 480-byte warp tile the span lies in (tile * AKF_WARPS + warp)
    int32_t span;        // 16, or 32: both chunks of a bit-parallel lane in one walk (the second chunk's info word is
    int32_t pad_;        // 0xC0000000 | index: "continued", no bytes of its own)
    uint8_t bytes[AK_SLOW_BYTES];      // the span's output when it fits (else the second pass walks again)
};

struct AkNfWork {
    uint32_t* info;            // [n_tiles * AK_BLOCK] per lane: emit mask, or 0x80000000 | work-list index
    int32_t* tile_total;       // [n_tiles * AKF_WARPS] output bytes of every 480-byte warp tile
    int64_t* tile_base;        // [n_tiles * AKF_WARPS + 1] exclusive prefix
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


#ifndef AKN3_MINB
#define AKN3_MINB 8
#endif
__global__ void __launch_bounds__(AKN3_THREADS, AKN3_MINB) ak_nf3_classify_kernel(const AkFastNormArgs A, const AkNfWork W) {
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
                    e.tile = tile * AKF_WARPS + 2 * warp + (lane >> 4);
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
        // the warp's 960 bytes are two of the writer's 480-byte warp tiles: lanes 1-15 and lanes 16-30 (the halo lanes count 0)
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        if ((lane & 15) == 0) W.tile_total[(size_t)tile * AKF_WARPS + 2 * warp + (lane >> 4)] = cnt;
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
        const size_t cta_tile = (size_t)e.tile / AKF_WARPS;
        const int64_t r0 = A.tile_row[cta_tile * AKF_WARPS], r1 = A.tile_row[(cta_tile + 1) * AKF_WARPS];
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

// ---- K1d: write the fast lanes' bytes (staged in shared memory, 16-byte stores) and the row offsets.  Warp-autonomous:
// a warp owns 480 text bytes (30 real lanes of 16 + the two halo slots of the info layout), its output base comes from
// the scan over the warp tiles' totals, its stage is its own -- no CTA barrier, no CTA scan.
#define AKF_WSTAGE (AKF_WARP_BYTES + 160)
__global__ void __launch_bounds__(AK_BLOCK) ak_nf_write_kernel(const AkFastNormArgs A, const AkNfWork W) {
    __shared__ __align__(16) uint8_t stage_all[AKF_WARPS][AKF_WSTAGE + 32];
    const AkBatch& B = A.B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* stage = stage_all[warp];
    const long long n_wt = (long long)B.n_tiles * AKF_WARPS;
    for (long long wt = (long long)blockIdx.x * AKF_WARPS + warp; wt < n_wt; wt += (long long)gridDim.x * AKF_WARPS) {
        const int64_t ws0 = A.base0 + wt * AKF_WARP_BYTES;
        const int64_t r0 = A.tile_row[wt], r1 = A.tile_row[wt + 1];
        const int64_t cs = ws0 + (int64_t)(lane - 1) * 16;
        AkChunk c;
        akf_load_lane(B, cs, c);
        const uint32_t info = W.info[(size_t)wt * 32 + lane];
        const bool slow = (info & 0x80000000u) != 0;
        const bool cont = slow && (info & 0x40000000u);          // second chunk of a 32-byte slow span: nothing of its own
        const unsigned int sidx = info & 0x3FFFFFFFu;
        int cnt = 0;
        if (slow) { if (sidx < W.slow_cap && !cont) cnt = W.slow[sidx].cnt; }
        else cnt = __popc(info);
        int inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += y;
        }
        const int total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        const int pre = inc - cnt;
        const int64_t base = W.tile_base[wt];
        const bool fits = base + total <= A.out_cap;
        const bool staged = fits && total <= AKF_WSTAGE;
        const int pad = (int)((uintptr_t)(A.out + base) & 15);
        if (!fits && lane == 0 && total > 0) ak_raise(B.result, AK_ST_OVERFLOW);
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
        __syncwarp();
        if (staged) {
            // stage[pad .. pad + total) -> out[base ..): stage and global share their alignment modulo 16.  The holes
            // of slow chunks are copied as garbage here and filled by the slow write kernel afterwards.
            uint8_t* g = A.out + base;
            int head = (16 - pad) & 15;
            if (head > total) head = total;
            if (lane < head) g[lane] = stage[pad + lane];
            const int body = (total - head) >> 4;
            for (int i = lane; i < body; i += 32)
                *reinterpret_cast<uint4*>(g + head + 16 * i) = *reinterpret_cast<const uint4*>(stage + pad + head + 16 * i);
            const int tail0 = head + (body << 4);
            if (lane < total - tail0) g[tail0 + lane] = stage[pad + tail0 + lane];
        }
        // row offsets of the rows that start in a fast chunk of this warp tile (slow chunks write their own)
        const int64_t r_end = r1 <= B.n_rows ? r1 : B.n_rows + 1;
        for (int64_t rr = r0; rr < r_end; rr += 32) {
            const int64_t r = rr + lane;
            const bool have = r < r_end;
            const int rel = have ? (int)(B.off[r] - ws0) : 0;
            const int th = 1 + (rel >> 4), i = rel & 15;
            const uint32_t e = __shfl_sync(0xFFFFFFFFu, info, th);
            const int p = __shfl_sync(0xFFFFFFFFu, pre, th);
            if (!have) continue;
            if (!(e & 0x80000000u)) A.out_off[r] = base + p + __popc(e & ((1u << i) - 1u));
            else {
                // the slow pass left it relative to the span's output; a continued chunk's prefix already includes the span
                int64_t adj = 0;
                if ((e & 0x40000000u) && (e & 0x3FFFFFFFu) < W.slow_cap) adj = W.slow[e & 0x3FFFFFFFu].cnt;
                A.out_off[r] += base + p - adj;
            }
        }
        __syncwarp();
    }
}
