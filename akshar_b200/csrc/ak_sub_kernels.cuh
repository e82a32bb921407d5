// Row walker kernels: the Unigram encoder for models the word-wise path cannot take (ak_tok_host.h ak_uni_wordwise) or
// AKSHAR_MODE_ROWS, and roman_phonetic_signature.  The encoders' fast path is ak_tok_kernels.cuh.
#pragma once
struct AkUniArgs {
    AkBatch B;
    AkUniDev U;
    uint32_t* back;            // scratch: row r uses back[(off[r] - text_begin) + 2 r ...]
    int64_t back_cap;          // entries the scratch holds
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
            const int64_t at = (rs - B.text_begin) + 2 * r;
            if (rs < B.text_begin || re < rs || re > B.text_end || at + (re - rs) + 2 > A.back_cap) {
                // offsets that do not describe rows of this text, or a scratch that is too small: say so instead of walking
                ak_raise(B.result, AK_ST_INTERNAL);
            } else {
                back = A.back + at;
                n = ak_unigram_forward(A.U, B.text, rs, re, back);
                cnt = ak_unigram_backtrack(A.U, back, n, nullptr, 0, 0);
            }
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
            if (back) ak_unigram_backtrack(A.U, back, n, A.ids, obase + cnt, A.id_cap);
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

