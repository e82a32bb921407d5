// K2 + K3 kernels: grapheme clusters and script runs (reference segment.py:40-201).
#pragma once
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



// ---- AKSHAR_SEG_MASK: the same boundaries as bit masks, one bit per text byte -------------------------------------------
// Bit (p - text_begin) of a mask is set when a cluster / script run ENDS at byte position p (text_begin < p <= text_end;
// the end of a non-empty row counts).  No compaction, so no counts, scans, temporary streams or copies: one pass, every
// lane stores its 32 bits.  1/8 byte of output per text byte and stream instead of 4 bytes per boundary -- what the host
// link carries in the normalize -> akshars -> script runs pipeline.  The tag of the run that ends at a bit is in two
// more planes: (t1 t0) = 00 devanagari, 01 roman, 10 other, 11 none (a row of digits / punctuation only).
struct AkSegMaskArgs {
    AkBatch B;
    AkTables T;
    uint32_t flags;
    const int64_t* wrow;               // first row at or after base0 + 480 k
    int64_t base0;
    uint32_t* cmask;
    uint32_t* rmask;
    uint32_t* t0;
    uint32_t* t1;
    int64_t n_words;
    int shift;                         // (text address) & 15: 0 = lanes and mask words coincide
};

#define AKSM_THREADS 128
__global__ void __launch_bounds__(AKSM_THREADS, 8) ak_seg_mask_kernel(const AkSegMaskArgs A) {
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    const bool want_c = (A.flags & AK_SEG_CLUSTERS) != 0, want_r = (A.flags & AK_SEG_RUNS) != 0;
    const bool matras = (A.flags & AK_SEG_MATRAS) != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tb = B.text_begin, te = B.text_end;
    const long long n_wt = (te - A.base0 + AKN3_WARP_BYTES) / AKN3_WARP_BYTES;
    uint32_t st = 0;
    long long n_c = 0, n_r = 0;
    for (long long wt = (long long)blockIdx.x * (AKSM_THREADS / 32) + warp; wt < n_wt; wt += (long long)gridDim.x * (AKSM_THREADS / 32)) {
        const int64_t ws = A.base0 + wt * AKN3_WARP_BYTES;
        const int64_t r_w0 = A.wrow[2 * wt], r_w2 = A.wrow[2 * wt + 2];
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
        const int64_t ss = cs < tb ? tb : cs;
        const int64_t se = cs + 32 > te + 1 ? te + 1 : cs + 32;
        const bool active = lane >= 1 && ss < se;                   // lane 31 too: its low bits complete lane 30's word when shift != 0
        const uint32_t tb_bit = (tb >= cs && tb < cs + 32) ? 1u << (int)(tb - cs) : 0u;
        uint32_t mc = 0, mr = 0, t0 = 0, t1 = 0;
        if (active && (lane <= 30 || A.shift)) {
            if (aks3_phase3(L, up2p, tb_bit, matras, want_c, want_r)) {
                const uint32_t rows_ev = L.rows & ~tb_bit;
                if (want_c) mc = L.brk | rows_ev;
                if (want_r) {
                    mr = L.rchg | rows_ev;
                    t0 = mr & ~(L.PD | L.PO);
                    t1 = mr & ~(L.PD | L.PR);
                }
            } else {
                uint32_t m4[4] = {0u, 0u, 0u, 0u};
                AkSegOut o;
                o.cluster_ends = nullptr; o.cluster_splits = nullptr; o.run_ends = nullptr; o.run_tags = nullptr; o.run_splits = nullptr;
                o.cbase = o.rbase = 0; o.ccap = o.rcap = 0;
                o.lane_masks = m4;
                o.lane_base = cs;
                const int64_t rlo = r_w0 > 0 ? r_w0 - 1 : 0, rhi = r_w2 > B.n_rows ? B.n_rows : r_w2;
                int64_t a, b;
                ak_seg_span(A.T, B.text, B.off, B.n_rows, rlo, lane == 31 ? B.n_rows : rhi, ss, se, A.flags, AK_LOOKBACK_LIMIT, true, o, a, b, st);
                mc = m4[0]; mr = m4[1]; t0 = m4[2]; t1 = m4[3];
            }
        }
        if (lane >= 1 && lane <= 30) { n_c += __popc(mc); n_r += __popc(mr); }
        // lane l writes mask word (wt * 30 + l - 1): its own bits from `shift` up, the next lane's below
        const int sh = A.shift;
        const uint32_t nc = __shfl_down_sync(0xFFFFFFFFu, mc, 1), nrm = __shfl_down_sync(0xFFFFFFFFu, mr, 1);
        const uint32_t n0 = __shfl_down_sync(0xFFFFFFFFu, t0, 1), n1 = __shfl_down_sync(0xFFFFFFFFu, t1, 1);
        const int64_t w = wt * 30 + (lane - 1);
        if (lane >= 1 && lane <= 30 && w < A.n_words) {
            if (want_c) A.cmask[w] = sh ? __funnelshift_r(mc, nc, sh) : mc;
            if (want_r) {
                A.rmask[w] = sh ? __funnelshift_r(mr, nrm, sh) : mr;
                A.t0[w] = sh ? __funnelshift_r(t0, n0, sh) : t0;
                A.t1[w] = sh ? __funnelshift_r(t1, n1, sh) : t1;
            }
        }
    }
    // totals: one atomic per warp at the end of its grid-stride walk
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        n_c += __shfl_xor_sync(0xFFFFFFFFu, n_c, d);
        n_r += __shfl_xor_sync(0xFFFFFFFFu, n_r, d);
    }
    if (lane == 0) {
        if (n_c) atomicAdd((unsigned long long*)&B.result[0], (unsigned long long)n_c);
        if (n_r) atomicAdd((unsigned long long*)&B.result[1], (unsigned long long)n_r);
    }
    ak_raise(B.result, st);
}
