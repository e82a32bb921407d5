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
// K2 + K3 from parallel bit streams (ak_seg3.cuh), offset form: END byte offsets of the clusters / runs of every row.
// Two passes over the text, both warp-autonomous (a warp owns 960 text bytes: 30 real lanes + 2 halo lanes): the count pass
// leaves the number of cluster ends and run ends of every warp tile (popcounts of the event masks; the exact walker counts
// for its slow lanes), two scans (ak_scan_counts_kernel) turn them into bases, the emit pass classifies again and writes
// every offset, tag and row split at its final place.  No temporary stream, no copy kernel, no ordering between warps
// (round 1 wrote to per-CTA slices of a temporary stream and copied: twice the output traffic, 1.5 ms per GiB of copies).
// ------------------------------------------------------------------------------------------------
struct AkSegOffArgs {
    AkBatch B;
    AkTables T;
    uint32_t flags;
    const int64_t* wrow;               // first row at or after base0 + 480 k
    int64_t base0;
    int32_t* c_count;                  // [n_wt]
    int32_t* r_count;
    const int64_t* c_base;             // [n_wt] exclusive prefixes
    const int64_t* r_base;
    AkSegOut o;                        // the final outputs
};

#define AKSO_THREADS 128
template <bool EMIT>
__global__ void __launch_bounds__(AKSO_THREADS, 8) ak_seg_off_kernel(const AkSegOffArgs A) {
    AkBatch B = A.B;
    if (!ak_batch_begin(B)) return;
    const bool want_c = (A.flags & AK_SEG_CLUSTERS) != 0, want_r = (A.flags & AK_SEG_RUNS) != 0;
    const bool matras = (A.flags & AK_SEG_MATRAS) != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tb = B.text_begin, te = B.text_end;
    const long long n_wt = (te - A.base0 + AKN3_WARP_BYTES) / AKN3_WARP_BYTES;
    uint32_t st = 0;
    for (long long wt = (long long)blockIdx.x * (AKSO_THREADS / 32) + warp; wt < n_wt; wt += (long long)gridDim.x * (AKSO_THREADS / 32)) {
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
        const bool real = lane >= 1 && lane <= 30;
        const int64_t ss = cs < tb ? tb : cs;
        const int64_t se = cs + 32 > te + 1 ? te + 1 : cs + 32;
        const bool active = real && ss < se;
        // index of the first row that starts at or after this lane's first position
        int64_t nr = 0;
        if (EMIT) {
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
        const int64_t rlo = r_w0 > 0 ? r_w0 - 1 : 0, rhi = r_w2 > B.n_rows ? B.n_rows : r_w2;
        bool slow = false;
        int cc = 0, rc = 0;
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
        int inc2 = cc | (rc << 16);
        const int mine2 = inc2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, inc2, d);
            if (lane >= d) inc2 += y;
        }
        if (!EMIT) {
            if (lane == 31) {
                A.c_count[wt] = inc2 & 0xFFFF;
                A.r_count[wt] = inc2 >> 16;
            }
            continue;
        }
        if (!active) continue;
        const int64_t cat = (want_c ? A.c_base[wt] : 0) + ((inc2 - mine2) & 0xFFFF);       // my first cluster end / run end
        const int64_t rat = (want_r ? A.r_base[wt] : 0) + ((inc2 - mine2) >> 16);
        const bool fits = cat + cc <= A.o.ccap && rat + rc <= A.o.rcap;
        if (!fits) st |= AK_ST_OVERFLOW;
        if (slow) {
            AkSegOut o = A.o;
            o.cbase = cat;
            o.rbase = rat;
            if (!fits) o.ccap = o.rcap = 0;
            uint32_t st2 = 0;
            int64_t a, b;
            ak_seg_span(A.T, B.text, B.off, B.n_rows, rlo, rhi, ss, se, A.flags, AK_LOOKBACK_LIMIT, true, o, a, b, st2);
        } else {
            const uint32_t mc = want_c ? (L.brk | rows_ev) : 0u, mr = want_r ? (L.rchg | rows_ev) : 0u;
            if (fits) {
                const int64_t rs_in = nr > 0 ? B.off[nr - 1] : B.off[0];
                if (want_c) aks3_emit(L, mc, cs, rs_in, A.o.cluster_ends + cat, nullptr);
                if (want_r) aks3_emit(L, mr, cs, rs_in, A.o.run_ends + rat, A.o.run_tags + rat);
            }
            if (L.rows)
                aks3_splits(L, mc, mr, cs, B.off, B.n_rows, nr, cat, rat, want_c ? A.o.cluster_splits : nullptr,
                            want_r ? A.o.run_splits : nullptr);
        }
    }
    ak_raise(B.result, st);
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
#ifndef AKSM_MINB
#define AKSM_MINB 8
#endif
__global__ void __launch_bounds__(AKSM_THREADS, AKSM_MINB) ak_seg_mask_kernel(const AkSegMaskArgs A) {
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
