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

#include <exception>
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
#include "ak_tok.cuh"
#include "ak_tok_host.h"
#include "ak_decode_host.h"
#include "unicode_tables.inc"


#include "ak_common.cuh"
#include "ak_norm_kernels.cuh"
#include "ak_seg_kernels.cuh"
#include "ak_sub_kernels.cuh"
#include "ak_tok_kernels.cuh"
#include "ak_wtok_kernels.cuh"
#include "ak_decode_kernels.cuh"
#include "ak_feat_kernels.cuh"
#include "ak_lines_kernels.cuh"

// ================================================================================================
// host side: context, model upload, C ABI
// ================================================================================================
struct AkcTable {
    AkWordCache work{};
    unsigned long long* image = nullptr;
    size_t bytes = 0;
    bool reset = false;            // restore the image at the next call whatever the fill
};

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
    // event-stream encoders (ak_tok.cuh): one word cache per model, image built at load
    AkcTable tok_cache[2];
    AkDecTable dec[2][2] = {};     // [kind][form]: piece tables of the on-device decode / detokenize (ak_decode.cuh)
    bool uni_fast = false;         // the Unigram model has the shape the word-wise path needs
    int occ_words[2] = {0, 0}, occ_resolve[2] = {0, 0}, occ_check = 0, occ_emit = 0;
    // optional CUDA-event timing of the dominant kernel of each stage (bench.py's roofline line)
    bool timing = false;
    bool wc_hold = false;          // akshar_word_cache_hold(1): never restore the image, however full the cache
    cudaEvent_t tev[AKSHAR_TIMER_COUNT][2] = {};
    bool tev_valid[AKSHAR_TIMER_COUNT] = {};
    int occ_norm = 0, occ_seg = 0, occ_uni = 0, occ_sig = 0, occ_nf_write = 0, occ_nf3 = 0, occ_sf3 = 0;
};

// the entry points run on the context's device and leave the caller's current device as they found it
struct AkDeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit AkDeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~AkDeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
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
    AkDeviceGuard device_guard(device);
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
    UP(kmap_keys, ak_tbl_kmap_keys, AK_N_KMAP, uint32_t)
    UP(kmap_off, ak_tbl_kmap_off, AK_N_KMAP + 1, uint16_t)
    UP(kmap_data, ak_tbl_kmap_data, AK_N_KMAP_DATA, uint32_t)
    UP(hf_unknown, ak_tbl_hf_unknown, AK_N_HF_UNKNOWN, uint32_t)
#undef UP
    ctx->T.n_decomp = AK_N_DECOMP;
    ctx->T.n_pairs = AK_N_PAIRS;
    ctx->T.n_ll = AK_N_LATIN_LOWER;
    ctx->T.n_fl = AK_N_FULL_LOWER;
    ctx->T.n_kmap = AK_N_KMAP;
    ctx->T.n_hf_unknown = AK_N_HF_UNKNOWN;
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_norm, ak_normalize_kernel, AK_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_nf_write, ak_nf_write_kernel, AK_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_nf3, ak_nf3_classify_kernel, AKN3_THREADS, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_sf3, ak_sf3_kernel, AKS3_THREADS, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_seg, ak_segment_kernel, AK_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_uni, ak_unigram_kernel, AK_ROWS_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_sig, ak_signature_kernel, AK_ROWS_BLOCK, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_words[0], ak_words_kernel<0>, AKW_THREADS, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_words[1], ak_words_kernel<1>, AKW_THREADS, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_resolve[0], ak_resolve_kernel<0>, AKR_THREADS, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_resolve[1], ak_resolve_kernel<1>, AKR_THREADS, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_check, ak_unicheck_kernel, AKL_THREADS, 0));
    AK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_emit, ak_emit_kernel<int32_t>, AKL_THREADS, 0));
    return AKSHAR_OK;
}

static void ak_free_list(std::vector<void*>& v) {
    for (void* p : v) cudaFree(p);
    v.clear();
}

void akshar_ctx_destroy(akshar_ctx* ctx) {
    if (!ctx) return;
    AkDeviceGuard device_guard(ctx->device);
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
    if (hold < 0) {
        ctx->tok_cache[0].reset = ctx->tok_cache[1].reset = true;
        return AKSHAR_OK;
    }
    ctx->wc_hold = hold != 0;
    return AKSHAR_OK;
}

int akshar_timing_enable(akshar_ctx* ctx, int enable) {
    AkDeviceGuard device_guard(ctx ? ctx->device : 0);
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
struct AkTokWs {
    size_t wrow, count, wt_ids, wt_base, wt_seg, wt_segx, scan_state, row_flag, row_ev, row_fix, pool, longpool, slots, total;
    size_t pool_ints, longpool_ints;
    int64_t n_wt;
};
#define AKT_CAP_MIN 256            // event slots per 960 text bytes with the minimum workspace (0.27 events per byte)
#define AKT_CAP_MAX 2048           // a row start and a word start at every byte: 1920
#define AKT_SLOT_BYTES 20          // event + resolved record + aux word
static AkTokWs ak_tok_ws(int64_t n_bytes, int64_t n_rows) {
    AkTokWs W;
    W.n_wt = n_bytes / AKT_WARP_BYTES + 3;
    size_t at = 0;
    W.wrow = at;       at += ak_align(((size_t)W.n_wt * 2 + 4) * 8);
    W.count = at;      at += ak_align((size_t)W.n_wt * 4);
    W.wt_ids = at;     at += ak_align((size_t)W.n_wt * 4);
    W.wt_base = at;    at += ak_align((size_t)W.n_wt * 8);
    W.wt_seg = at;     at += ak_align((size_t)W.n_wt * 8);
    W.wt_segx = at;    at += ak_align((size_t)W.n_wt * 4);
    W.scan_state = at; at += 2 * ak_align(((size_t)W.n_wt / AKS_TILE + 2) * 8);
    W.row_flag = at;   at += ak_align((size_t)n_rows + 1);
    W.row_ev = at;     at += ak_align(((size_t)n_rows + 2) * 4);
    W.row_fix = at;    at += ak_align(((size_t)n_rows + 1) * 8);
    W.pool_ints = (size_t)n_bytes / 4 + 65536;
    W.pool = at;       at += ak_align(W.pool_ints * 4);
    W.longpool_ints = (size_t)n_bytes / 8 + 65536;
    W.longpool = at;   at += ak_align(W.longpool_ints * 4);
    W.slots = at;      // events, resolved records, aux words: sized at run time from the real workspace
    at += (size_t)W.n_wt * AKT_CAP_MIN * (AKT_SLOT_BYTES + 1) + 8192;
    W.total = at;
    return W;
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
    const size_t tok = ak_tok_ws(n_bytes, n_rows).total;
    if (tok > m) m = tok;
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
    // AKSHAR_DEBUG_SYNC=1: wait for every kernel and say which one faulted (debugging aid; never set in production)
    static const bool debug_sync = getenv("AKSHAR_DEBUG_SYNC") != nullptr;
    if (e == cudaSuccess && debug_sync) {
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) fprintf(stderr, "akshar_b200: kernel '%s' failed: %s\n", what, cudaGetErrorString(e));
    }
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
        // tile totals -> tile bases + the output length (single pass, look-back over tiles of 4096 entries; the state words
        // are the second quarter of the zeroed state area, the ticket is control word 6)
        ak_scan_counts_kernel<<<ak_grid(ctx, 4, F.B.n_tiles / AKS_TILE + 1), AKS_THREADS, 0, C.stream>>>(
            W.tile_total, (long long)F.B.n_tiles, nullptr, 1, W.tile_base, B.totals, (int*)C.ws + 6, C.B.state1,
            (unsigned int*)&B.result[2]);
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
    AkDeviceGuard device_guard(ctx ? ctx->device : 0);
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

// AKSHAR_SEG_MASK launch: warp-tile row search + the one-pass mask kernel.  max_bytes bounds the text when its length is
// only known on the device (B.dyn_end); the planes t0 / t1 are n_words apart
static int ak_run_seg_mask(akshar_ctx* ctx, AkCall& C, const AkBatch& B, int64_t max_bytes, uint32_t flags, uint32_t* cmask,
                           uint32_t* rmask, uint32_t* tags, int64_t n_words) {
    int rc;
    AkSegMaskArgs M;
    M.B = B;
    M.T = ctx->T;
    M.flags = flags;
    M.shift = (int)(((uintptr_t)B.text + (uintptr_t)B.text_begin) & 15u);
    M.base0 = B.text_begin - M.shift;
    M.cmask = cmask;
    M.rmask = rmask;
    M.t0 = tags;
    M.t1 = tags ? tags + n_words : nullptr;
    M.n_words = n_words;
    const int64_t n_wt = (B.text_begin + max_bytes - M.base0 + AKN3_WARP_BYTES) / AKN3_WARP_BYTES;
    int64_t* wrow = (int64_t*)(C.ws + C.L.scratch);
    M.wrow = wrow;
    const int entries = (int)(n_wt * 2 + 3);
    ak_warp_rows_kernel<<<(entries + 255) / 256, 256, 0, C.stream>>>(B, M.base0, entries, wrow);
    if ((rc = ak_after_launch(ctx, "segment-warp-rows"))) return rc;
    {
        AkTimed tm(ctx, AKSHAR_TIMER_SEGMENT, C.stream);
        ak_seg_mask_kernel<<<ak_grid(ctx, 8, (int)((n_wt + AKSM_THREADS / 32 - 1) / (AKSM_THREADS / 32))), AKSM_THREADS, 0, C.stream>>>(M);
    }
    return ak_after_launch(ctx, "segment-mask");
}

int akshar_segment_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                         int64_t text_begin, int64_t text_end, uint32_t flags, int mode, int32_t* d_cluster_ends,
                         int64_t cluster_capacity, int64_t* d_cluster_splits, int32_t* d_run_ends, uint8_t* d_run_tags,
                         int64_t run_capacity, int64_t* d_run_splits, int64_t* d_result, void* d_workspace,
                         size_t workspace_bytes, void* stream) {
    AkDeviceGuard device_guard(ctx ? ctx->device : 0);
    AkCall C;
    int rc = ak_begin(ctx, d_text, d_row_offsets, n_rows, text_begin, text_end, mode, 0, d_result, d_workspace, workspace_bytes,
                      stream, C);
    if (rc) return rc;
    const bool want_c = (flags & AKSHAR_SEG_CLUSTERS) != 0, want_r = (flags & AKSHAR_SEG_RUNS) != 0;
    if (flags & AKSHAR_SEG_MASK) {
        // boundaries as bit masks: d_cluster_ends / d_run_ends receive mask words, d_run_tags the two tag planes
        const int64_t n_words = (text_end - text_begin + 32) / 32;
        if ((flags & ~15u) || (!want_c && !want_r) || ((flags & AKSHAR_SEG_MATRAS) && !want_c) || mode != AKSHAR_MODE_TILES ||
            (want_c && (!d_cluster_ends || cluster_capacity < n_words)) ||
            (want_r && (!d_run_ends || !d_run_tags || run_capacity < n_words))) {
            ctx->err = "bad argument (AKSHAR_SEG_MASK: tile mode, capacities in 32-bit words >= (bytes + 32) / 32)";
            return AKSHAR_E_ARG;
        }
        return ak_run_seg_mask(ctx, C, C.B, text_end - text_begin, flags & 7u, (uint32_t*)d_cluster_ends, (uint32_t*)d_run_ends,
                               (uint32_t*)d_run_tags, n_words);
    }
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
        const int grid = ak_grid(ctx, ctx->occ_sf3, ((nwt + 1) / 2 + AKS3_THREADS / 32 - 1) / (AKS3_THREADS / 32));
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
            ak_sf3_kernel<<<grid, AKS3_THREADS, 0, C.stream>>>(F);
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

int akshar_word_tokenize_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                               int64_t text_begin, int64_t text_end, int rule, int32_t* d_word_begin, int32_t* d_word_end,
                               int64_t word_capacity, int64_t* d_word_splits, uint8_t* d_row_flags, int64_t* d_result,
                               void* d_workspace, size_t workspace_bytes, void* stream) {
    AkDeviceGuard device_guard(ctx ? ctx->device : 0);
    AkCall C;
    int rc = ak_begin(ctx, d_text, d_row_offsets, n_rows, text_begin, text_end, AKSHAR_MODE_TILES, 0, d_result, d_workspace,
                      workspace_bytes, stream, C);
    if (rc) return rc;
    if ((rule != AKSHAR_WORDS_HINDI && rule != AKSHAR_WORDS_SPLIT) || !d_word_splits || word_capacity < 0 ||
        ((!d_word_begin || !d_word_end) && word_capacity > 0)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (n_rows == 0) return ak_empty_rows(ctx, d_word_splits, nullptr, C.stream);
    if (d_row_flags) AK_CUDA(ctx, cudaMemsetAsync(d_row_flags, 0, (size_t)n_rows, C.stream));
    AkWtArgs A;
    A.B = C.B;
    A.base0 = text_begin - (int64_t)(((uintptr_t)d_text + (uintptr_t)text_begin) & 15u);
    const int64_t n_wt = (text_end - A.base0 + AKT_WARP_BYTES) / AKT_WARP_BYTES;
    char* wp = C.ws + C.L.scratch;
    int64_t* wrow = (int64_t*)wp;       wp += ak_align(((size_t)n_wt * 2 + 3) * 8);
    A.count = (int32_t*)wp;             wp += ak_align((size_t)n_wt * 4);
    int64_t* wt_base = (int64_t*)wp;
    A.wrow = wrow;
    A.base = wt_base;
    A.mode = rule == AKSHAR_WORDS_HINDI ? AKW_MODE_HINDI : AKW_MODE_SPLIT;
    A.begin = d_word_begin;
    A.end = d_word_end;
    A.cap = word_capacity;
    A.splits = d_word_splits;
    A.row_flags = d_row_flags;
    const int entries = (int)(n_wt * 2 + 3);
    ak_warp_rows_kernel<<<(entries + 255) / 256, 256, 0, C.stream>>>(C.B, A.base0, entries, wrow);
    if ((rc = ak_after_launch(ctx, "words-warp-rows"))) return rc;
    const int grid = ak_grid(ctx, 8, (int)((n_wt + AKWT_THREADS / 32 - 1) / (AKWT_THREADS / 32)));
    ak_wtok_kernel<false><<<grid, AKWT_THREADS, 0, C.stream>>>(A);
    if ((rc = ak_after_launch(ctx, "words-count"))) return rc;
    ak_scan_counts_kernel<<<ak_grid(ctx, 4, (int)(n_wt / AKS_TILE + 1)), AKS_THREADS, 0, C.stream>>>(
        A.count, (long long)n_wt, nullptr, 1, wt_base, d_result, (int*)C.ws + 5, C.B.state0, (unsigned int*)&d_result[2]);
    if ((rc = ak_after_launch(ctx, "words-scan"))) return rc;
    {
        AkTimed tm(ctx, AKSHAR_TIMER_WORDTOK, C.stream);
        ak_wtok_kernel<true><<<grid, AKWT_THREADS, 0, C.stream>>>(A);
    }
    return ak_after_launch(ctx, "words-emit");
}

}  // extern "C"

// ---- ids -> text ----------------------------------------------------------------------------------------------------
struct AkDecWs {
    size_t mark, count, base, tpre, state, total;
};
static AkDecWs ak_dec_ws(int64_t n_ids) {
    AkDecWs W;
    const size_t tiles = (size_t)(n_ids / AKD_TILE + 2);
    W.mark = 256;
    W.count = W.mark + ak_align((size_t)n_ids + 16);
    W.base = W.count + ak_align(tiles * 4);
    W.tpre = W.base + ak_align(tiles * 8);
    W.state = W.tpre + ak_align(((size_t)n_ids / AKD_PER + 2) * 4);
    W.total = W.state + ak_align((tiles / AKS_TILE + 2) * 8);
    return W;
}

template <class IdT>
static int ak_run_decode(akshar_ctx* ctx, const AkDecTable& D, int form, const void* d_ids, int64_t n_ids, const int64_t* d_row_splits,
                         int64_t n_rows, uint8_t* d_out_text, int64_t out_capacity, int64_t* d_out_row_offsets, int64_t* d_result,
                         char* ws, cudaStream_t s) {
    const AkDecWs W = ak_dec_ws(n_ids);
    int rc;
    AK_CUDA(ctx, cudaMemsetAsync(ws, 0, W.count, s));                          // control block + marks
    AK_CUDA(ctx, cudaMemsetAsync(ws + W.state, 0, W.total - W.state, s));
    AK_CUDA(ctx, cudaMemsetAsync(d_result, 0, 4 * sizeof(int64_t), s));
    AkDecArgs<IdT> A;
    A.D = D;
    A.form = form;
    A.ids = (const IdT*)d_ids;
    A.n_ids = n_ids;
    A.splits = d_row_splits;
    A.n_rows = n_rows;
    A.mark = (uint8_t*)(ws + W.mark);
    A.count = (int32_t*)(ws + W.count);
    A.base = (const int64_t*)(ws + W.base);
    A.tpre = (int32_t*)(ws + W.tpre);
    A.out = d_out_text;
    A.cap = out_capacity;
    A.out_off = d_out_row_offsets;
    A.result = d_result;
    const int64_t n_tiles = (n_ids + AKD_TILE - 1) / AKD_TILE;
    ak_dec_mark_kernel<IdT><<<ak_grid(ctx, 8, (int)((n_rows + 255) / 256)), 256, 0, s>>>(A);
    if ((rc = ak_after_launch(ctx, "decode-mark"))) return rc;
    const int grid = ak_grid(ctx, 8, (int)n_tiles);
    ak_dec_kernel<IdT, false><<<grid, AKD_THREADS, 0, s>>>(A);
    if ((rc = ak_after_launch(ctx, "decode-count"))) return rc;
    ak_scan_counts_kernel<<<ak_grid(ctx, 4, (int)(n_tiles / AKS_TILE + 1)), AKS_THREADS, 0, s>>>(
        A.count, (long long)n_tiles, nullptr, 1, (int64_t*)(ws + W.base), d_result, (int*)ws, (unsigned long long*)(ws + W.state),
        (unsigned int*)&d_result[2]);
    if ((rc = ak_after_launch(ctx, "decode-scan"))) return rc;
    {
        AkTimed tm(ctx, AKSHAR_TIMER_DECODE, s);
        ak_dec_kernel<IdT, true><<<grid, AKD_THREADS, 0, s>>>(A);
    }
    if ((rc = ak_after_launch(ctx, "decode-write"))) return rc;
    ak_dec_rowoff_kernel<IdT><<<ak_grid(ctx, 8, (int)((n_rows + 256) / 256)), 256, 0, s>>>(A);
    return ak_after_launch(ctx, "decode-row-offsets");
}

extern "C" {

size_t akshar_decode_workspace_bytes(int64_t n_ids, int64_t n_rows) {
    if (n_ids < 0 || n_rows < 0) return 0;
    return ak_dec_ws(n_ids).total;
}

int akshar_decode_batch(akshar_ctx* ctx, int kind, int form, const void* d_ids, int ids_u16, int64_t n_ids, const int64_t* d_row_splits,
                        int64_t n_rows, uint8_t* d_out_text, int64_t out_capacity, int64_t* d_out_row_offsets, int64_t* d_result,
                        void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!ctx) return AKSHAR_E_ARG;
    AkDeviceGuard device_guard(ctx->device);
    if ((kind != 0 && kind != 1) || (form != AKSHAR_FORM_DECODE && form != AKSHAR_FORM_DETOKENIZE) || n_ids < 0 || n_rows < 0 ||
        (!d_ids && n_ids > 0) || !d_row_splits || !d_out_row_offsets || !d_result || out_capacity < 0 || (!d_out_text && out_capacity > 0)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (!(kind == 0 ? ctx->has_bpe : ctx->has_uni)) {
        ctx->err = "no model loaded";
        return AKSHAR_E_NOMODEL;
    }
    const size_t need = ak_dec_ws(n_ids).total;
    if (!d_workspace || workspace_bytes < need) {
        ctx->err = "workspace too small: need " + std::to_string(need) + " bytes";
        return AKSHAR_E_WORKSPACE;
    }
    const AkDecTable& D = ctx->dec[kind][form == AKSHAR_FORM_DECODE ? AKD_FORM_DECODE : AKD_FORM_DETOK];
    const int f = form == AKSHAR_FORM_DECODE ? AKD_FORM_DECODE : AKD_FORM_DETOK;
    if (ids_u16)
        return ak_run_decode<uint16_t>(ctx, D, f, d_ids, n_ids, d_row_splits, n_rows, d_out_text, out_capacity, d_out_row_offsets,
                                       d_result, (char*)d_workspace, (cudaStream_t)stream);
    return ak_run_decode<int32_t>(ctx, D, f, d_ids, n_ids, d_row_splits, n_rows, d_out_text, out_capacity, d_out_row_offsets, d_result,
                                  (char*)d_workspace, (cudaStream_t)stream);
}

// ---- per-sentence statistics and cluster merging over the segment kernel's outputs -----------------------------------
int akshar_composition_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                             const int64_t* d_cluster_splits, const int32_t* d_run_ends, const uint8_t* d_run_tags,
                             const int64_t* d_run_splits, int32_t* d_stats, void* stream) {
    if (!ctx) return AKSHAR_E_ARG;
    AkDeviceGuard device_guard(ctx->device);
    if (!d_row_offsets || n_rows < 0 || !d_cluster_splits || !d_run_splits || (n_rows > 0 && !d_stats)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (n_rows == 0) return AKSHAR_OK;
    AkCompArgs A;
    A.text = d_text;
    A.off = d_row_offsets;
    A.n_rows = n_rows;
    A.cluster_splits = d_cluster_splits;
    A.run_ends = d_run_ends;
    A.run_tags = d_run_tags;
    A.run_splits = d_run_splits;
    A.stats = d_stats;
    ak_comp_kernel<<<ak_grid(ctx, 8, (int)((n_rows + 7) / 8)), 256, 0, (cudaStream_t)stream>>>(A);
    return ak_after_launch(ctx, "composition");
}

static size_t ak_cm_ws(int64_t n, size_t* flag, size_t* count, size_t* base, size_t* state) {
    const size_t tiles = (size_t)(n / AKCM_TILE + 2);
    *flag = 256;
    *count = *flag + ak_align((size_t)n + 16);
    *base = *count + ak_align(tiles * 4);
    *state = *base + ak_align(tiles * 8);
    return *state + ak_align((tiles / AKS_TILE + 2) * 8);
}

size_t akshar_merge_workspace_bytes(int64_t n_clusters) {
    size_t a, b, c, d;
    return n_clusters < 0 ? 0 : ak_cm_ws(n_clusters, &a, &b, &c, &d);
}

int akshar_merge_clusters_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                                const int32_t* d_cluster_ends, const int64_t* d_cluster_splits, int64_t n_clusters, int rule,
                                int32_t* d_out_ends, int64_t out_capacity, int64_t* d_out_splits, int64_t* d_result,
                                void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!ctx) return AKSHAR_E_ARG;
    AkDeviceGuard device_guard(ctx->device);
    if (!d_row_offsets || n_rows < 0 || n_clusters < 0 || !d_cluster_splits || !d_out_splits || !d_result || out_capacity < 0 ||
        (rule != AKSHAR_MERGE_AKSHARA && rule != AKSHAR_MERGE_NUKTA) || (n_clusters > 0 && (!d_cluster_ends || !d_text)) ||
        (out_capacity > 0 && !d_out_ends)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    size_t o_flag, o_count, o_base, o_state;
    const size_t need = ak_cm_ws(n_clusters, &o_flag, &o_count, &o_base, &o_state);
    if (!d_workspace || workspace_bytes < need) {
        ctx->err = "workspace too small: need " + std::to_string(need) + " bytes";
        return AKSHAR_E_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = (char*)d_workspace;
    int rc;
    AK_CUDA(ctx, cudaMemsetAsync(ws, 0, 256, s));
    AK_CUDA(ctx, cudaMemsetAsync(ws + o_state, 0, need - o_state, s));
    AK_CUDA(ctx, cudaMemsetAsync(d_result, 0, 4 * sizeof(int64_t), s));
    AkCmArgs A;
    A.text = d_text;
    A.off = d_row_offsets;
    A.n_rows = n_rows;
    A.ends = d_cluster_ends;
    A.splits = d_cluster_splits;
    A.n = n_clusters;
    A.rule = rule == AKSHAR_MERGE_AKSHARA ? AKCM_AKSHARA : AKCM_NUKTA;
    A.flag = (uint8_t*)(ws + o_flag);
    A.count = (int32_t*)(ws + o_count);
    A.base = (const int64_t*)(ws + o_base);
    A.out_ends = d_out_ends;
    A.out_splits = d_out_splits;
    A.cap = out_capacity;
    A.result = d_result;
    const int64_t n_tiles = (n_clusters + AKCM_TILE - 1) / AKCM_TILE;
    if (n_clusters > 0) {
        ak_cm_flag_kernel<<<ak_grid(ctx, 8, (int)((n_clusters + 255) / 256)), 256, 0, s>>>(A);
        if ((rc = ak_after_launch(ctx, "merge-flags"))) return rc;
        ak_cm_kernel<false><<<ak_grid(ctx, 8, (int)n_tiles), AKCM_THREADS, 0, s>>>(A);
        if ((rc = ak_after_launch(ctx, "merge-count"))) return rc;
    }
    ak_scan_counts_kernel<<<ak_grid(ctx, 4, (int)(n_tiles / AKS_TILE + 1)), AKS_THREADS, 0, s>>>(
        A.count, (long long)n_tiles, nullptr, 1, (int64_t*)(ws + o_base), d_result, (int*)ws, (unsigned long long*)(ws + o_state),
        (unsigned int*)&d_result[2]);
    if ((rc = ak_after_launch(ctx, "merge-scan"))) return rc;
    ak_cm_kernel<true><<<ak_grid(ctx, 8, (int)(n_tiles > 0 ? n_tiles : 1)), AKCM_THREADS, 0, s>>>(A);
    return ak_after_launch(ctx, "merge-write");
}

// ---- file bytes -> rows -----------------------------------------------------------------------------------------------
struct AkLinesWs {
    size_t fn, begin, end, len, state, total;
};
static AkLinesWs ak_lines_ws(int64_t n_bytes, int64_t row_capacity) {
    AkLinesWs W;
    const size_t tiles = (size_t)(n_bytes / AKLN_TILE + 2);
    W.fn = 256;
    W.begin = W.fn + ak_align(tiles * sizeof(AkLineFn));
    W.end = W.begin + ak_align(((size_t)row_capacity + 1) * 8);
    W.len = W.end + ak_align(((size_t)row_capacity + 1) * 8);
    W.state = W.len + ak_align(((size_t)row_capacity + 1) * 4);
    W.total = W.state + ak_align(((size_t)row_capacity / AKS_TILE + 2) * 8);
    return W;
}

size_t akshar_lines_workspace_bytes(int64_t n_bytes, int64_t row_capacity) {
    if (n_bytes < 0 || row_capacity < 0) return 0;
    return ak_lines_ws(n_bytes, row_capacity).total;
}

int akshar_lines_batch(akshar_ctx* ctx, const uint8_t* d_file, int64_t n_bytes, uint8_t* d_out_text, int64_t out_capacity,
                       int64_t* d_out_row_offsets, int64_t row_capacity, int64_t* d_result, void* d_workspace, size_t workspace_bytes,
                       void* stream) {
    if (!ctx) return AKSHAR_E_ARG;
    AkDeviceGuard device_guard(ctx->device);
    if (n_bytes < 0 || (!d_file && n_bytes > 0) || out_capacity < 0 || (!d_out_text && out_capacity > 0) || !d_out_row_offsets ||
        row_capacity < 0 || !d_result) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    const AkLinesWs W = ak_lines_ws(n_bytes, row_capacity);
    if (!d_workspace || workspace_bytes < W.total) {
        ctx->err = "workspace too small: need " + std::to_string(W.total) + " bytes";
        return AKSHAR_E_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = (char*)d_workspace;
    int rc;
    AK_CUDA(ctx, cudaMemsetAsync(ws, 0, 256, s));
    AK_CUDA(ctx, cudaMemsetAsync(ws + W.state, 0, W.total - W.state, s));
    AK_CUDA(ctx, cudaMemsetAsync(d_result, 0, 4 * sizeof(int64_t), s));
    AkLinesArgs A;
    A.text = d_file;
    A.n = n_bytes;
    A.n_tiles = n_bytes / AKLN_TILE + 1;
    A.tile_fn = (AkLineFn*)(ws + W.fn);
    A.begin = (int64_t*)(ws + W.begin);
    A.end = (int64_t*)(ws + W.end);
    A.cap = row_capacity;
    A.result = d_result;
    const int grid = ak_grid(ctx, 8, (int)((A.n_tiles + 7) / 8));
    ak_lines_kernel<false><<<grid, 256, 0, s>>>(A);
    if ((rc = ak_after_launch(ctx, "lines-summaries"))) return rc;
    ak_lines_resolve_kernel<<<1, 1024, 0, s>>>(A);
    if ((rc = ak_after_launch(ctx, "lines-resolve"))) return rc;
    {
        AkTimed tm(ctx, AKSHAR_TIMER_LINES, s);
        ak_lines_kernel<true><<<grid, 256, 0, s>>>(A);
    }
    if ((rc = ak_after_launch(ctx, "lines-emit"))) return rc;
    int32_t* len = (int32_t*)(ws + W.len);
    const int rgrid = ak_grid(ctx, 8, (int)((row_capacity + 255) / 256));
    ak_lines_len_kernel<<<rgrid, 256, 0, s>>>(A.begin, A.end, row_capacity, len, d_result);
    if ((rc = ak_after_launch(ctx, "lines-lengths"))) return rc;
    ak_scan_counts_kernel<<<ak_grid(ctx, 4, (int)(row_capacity / AKS_TILE + 1)), AKS_THREADS, 0, s>>>(
        len, 0, (const long long*)(d_result + 3), 1, d_out_row_offsets, d_result + 1, (int*)ws, (unsigned long long*)(ws + W.state),
        (unsigned int*)&d_result[2]);
    if ((rc = ak_after_launch(ctx, "lines-scan"))) return rc;
    ak_lines_gather_kernel<<<ak_grid(ctx, 8, (int)((row_capacity + 7) / 8)), 256, 0, s>>>(d_file, A.begin, len, d_out_row_offsets, d_out_text,
                                                                                         out_capacity, d_result);
    return ak_after_launch(ctx, "lines-gather");
}

int akshar_join_rows(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows, int sep, uint8_t* d_out,
                     void* stream) {
    if (!ctx) return AKSHAR_E_ARG;
    AkDeviceGuard device_guard(ctx->device);
    if (!d_row_offsets || n_rows < 0 || sep < 0 || sep > 255 || (n_rows > 0 && !d_out)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (n_rows == 0) return AKSHAR_OK;
    ak_join_rows_kernel<<<ak_grid(ctx, 8, (int)((n_rows + 7) / 8)), 256, 0, (cudaStream_t)stream>>>(d_text, d_row_offsets, n_rows, (uint8_t)sep, d_out);
    return ak_after_launch(ctx, "join-rows");
}

int akshar_normalize_segment_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                                   int64_t text_begin, int64_t text_end, uint32_t norm_flags, uint32_t seg_flags, uint8_t* d_norm_text,
                                   int64_t norm_capacity, int64_t* d_norm_row_offsets, uint32_t* d_cluster_mask, uint32_t* d_run_mask,
                                   uint32_t* d_run_tag_planes, int64_t mask_words, int64_t* d_result, void* d_workspace,
                                   size_t workspace_bytes, void* stream) {
    AkDeviceGuard device_guard(ctx ? ctx->device : 0);
    if (!ctx) return AKSHAR_E_ARG;
    const bool want_c = (seg_flags & AKSHAR_SEG_CLUSTERS) != 0, want_r = (seg_flags & AKSHAR_SEG_RUNS) != 0;
    const int64_t n_bytes = text_end - text_begin;
    const int64_t max_bytes = n_bytes > norm_capacity ? n_bytes : norm_capacity;
    if (norm_capacity < 0 || !d_norm_row_offsets || (!d_norm_text && norm_capacity > 0) || (norm_flags & ~15u) || (seg_flags & ~7u) ||
        (!want_c && !want_r) || ((seg_flags & AKSHAR_SEG_MATRAS) && !want_c) || mask_words < (norm_capacity + 32) / 32 ||
        (want_c && !d_cluster_mask) || (want_r && (!d_run_mask || !d_run_tag_planes))) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if (workspace_bytes < ak_ws_layout(max_bytes, n_rows).total) {
        ctx->err = "workspace too small: need " + std::to_string(ak_ws_layout(max_bytes, n_rows).total) + " bytes";
        return AKSHAR_E_WORKSPACE;
    }
    AkCall C;
    int rc = ak_begin(ctx, d_text, d_row_offsets, n_rows, text_begin, text_end, AKSHAR_MODE_TILES, 0, d_result, d_workspace,
                      workspace_bytes, stream, C);
    if (rc) return rc;
    C.L = ak_ws_layout(max_bytes, n_rows);
    const size_t tiles = (size_t)ak_tiles_of(max_bytes, n_rows);
    AK_CUDA(ctx, cudaMemsetAsync(C.ws, 0, 256 + ak_align(4 * tiles * 8), C.stream));
    C.B.state0 = (unsigned long long*)(C.ws + C.L.state);
    C.B.state1 = C.B.state0 + tiles;
    if (n_rows == 0) return ak_empty_rows(ctx, d_norm_row_offsets, nullptr, C.stream);
    // stage 1: normalize_text; its byte total lands in result[3]
    AkBatch B1 = C.B;
    B1.totals = d_result + 3;
    if ((rc = ak_run_normalize(ctx, C, B1, norm_flags, d_norm_text, norm_capacity, d_norm_row_offsets))) return rc;
    // stage 2: akshars / script runs of the normalized rows as bit masks; their length is only known on the device
    AkBatch B2 = C.B;
    B2.text = d_norm_text;
    B2.off = d_norm_row_offsets;
    B2.text_begin = 0;
    B2.text_end = 0;
    B2.dyn_end = d_result + 3;
    return ak_run_seg_mask(ctx, C, B2, norm_capacity, seg_flags, d_cluster_mask, d_run_mask, d_run_tag_planes, mask_words);
}

int akshar_signature_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                           int64_t text_begin, int64_t text_end, uint8_t* d_out_text, int64_t out_capacity,
                           int64_t* d_out_row_offsets, int64_t* d_result, void* d_workspace, size_t workspace_bytes,
                           void* stream) {
    AkDeviceGuard device_guard(ctx ? ctx->device : 0);
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

static int ak_load_bpe_json(akshar_ctx* ctx, const char* json, size_t len);
static int ak_load_spm_model(akshar_ctx* ctx, const void* proto, size_t len);

// no exception crosses the ABI: a model that makes a parser or an allocation throw is a model error
int akshar_load_bpe_json(akshar_ctx* ctx, const char* json, size_t len) {
    if (!ctx || !json) return AKSHAR_E_ARG;
    AkDeviceGuard device_guard(ctx->device);
    try {
        return ak_load_bpe_json(ctx, json, len);
    } catch (const std::exception& e) {
        ctx->err = std::string("tokenizer JSON: ") + e.what();
        return AKSHAR_E_MODEL;
    } catch (...) {
        ctx->err = "tokenizer JSON: unexpected failure";
        return AKSHAR_E_MODEL;
    }
}
int akshar_load_spm_model(akshar_ctx* ctx, const void* proto, size_t len) {
    if (!ctx || !proto) return AKSHAR_E_ARG;
    AkDeviceGuard device_guard(ctx->device);
    try {
        return ak_load_spm_model(ctx, proto, len);
    } catch (const std::exception& e) {
        ctx->err = std::string("SentencePiece model: ") + e.what();
        return AKSHAR_E_MODEL;
    } catch (...) {
        ctx->err = "SentencePiece model: unexpected failure";
        return AKSHAR_E_MODEL;
    }
}

}  // extern "C"

// the image goes to the device twice: the pristine copy and the working copy the kernels add to
static int ak_install_cache(akshar_ctx* ctx, std::vector<void*>& owner, const std::vector<unsigned long long>& img, AkcTable& t) {
    const unsigned long long* dimg = nullptr;
    int rc;
    if ((rc = ak_upload<unsigned long long>(ctx, owner, img.data(), img.size(), &dimg))) return rc;
    void* work = nullptr;
    AK_CUDA(ctx, cudaMalloc(&work, img.size() * 8));
    owner.push_back(work);
    AK_CUDA(ctx, cudaMemcpy(work, dimg, img.size() * 8, cudaMemcpyDeviceToDevice));
    void* counter = nullptr;
    AK_CUDA(ctx, cudaMalloc(&counter, 256));
    owner.push_back(counter);
    AK_CUDA(ctx, cudaMemset(counter, 0, 256));
    ctx->wc_hold = false;             // a new model always starts without a hold
    t.image = (unsigned long long*)dimg;
    t.work.e = (unsigned long long*)work;
    t.work.bits = AKC_BITS;
    t.work.inserted = (unsigned long long*)counter;
    t.bytes = img.size() * 8;
    t.reset = false;
    return AKSHAR_OK;
}

// piece tables of decode / detokenize (ak_decode_host.h) to the device
static int ak_install_decode(akshar_ctx* ctx, std::vector<void*>& owner, const AkDecHost& h, AkDecTable& t) {
    int rc;
    if ((rc = ak_upload<uint32_t>(ctx, owner, h.off.data(), h.off.size(), &t.off))) return rc;
    if ((rc = ak_upload<uint8_t>(ctx, owner, h.bytes.data(), h.bytes.size(), &t.bytes))) return rc;
    if ((rc = ak_upload<uint8_t>(ctx, owner, h.flags.data(), h.flags.size(), &t.flags))) return rc;
    t.size = (int32_t)h.flags.size();
    t.strict = h.strict;
    return AKSHAR_OK;
}

static int ak_load_bpe_json(akshar_ctx* ctx, const char* json, size_t len) {
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
    if ((rc = ak_upload<uint8_t>(ctx, ctx->bpe_allocs, h.sp_bytes.data(), h.sp_bytes.size(), &d.sp_bytes))) return rc;
    if ((rc = ak_upload<uint16_t>(ctx, ctx->bpe_allocs, h.sp_off.data(), h.sp_off.size(), &d.sp_off))) return rc;
    if ((rc = ak_upload<int32_t>(ctx, ctx->bpe_allocs, h.sp_ids.data(), h.sp_ids.size(), &d.sp_ids))) return rc;
    d.n_sp = (int)h.sp_ids.size();
    d.n_cp = (int)h.cp_keys.size();
    d.mbits = h.mbits;
    d.bos = h.bos;
    d.eos = h.eos;
    {
        AkTables ht{};
        ht.page_index = ak_tbl_page_index;
        ht.leaves = ak_tbl_leaves;
        const std::vector<unsigned long long> img = ak_build_bpe_image(h, ht, AKC_BITS);
        if ((rc = ak_install_cache(ctx, ctx->bpe_allocs, img, ctx->tok_cache[0]))) return rc;
    }
    for (int form = 0; form < 2; ++form)
        if ((rc = ak_install_decode(ctx, ctx->bpe_allocs, ak_build_bpe_decode(h, form), ctx->dec[0][form]))) return rc;
    ctx->bpe_d = d;
    ctx->bpe_h = std::move(h);
    ctx->has_bpe = true;
    return AKSHAR_OK;
}

static int ak_load_spm_model(akshar_ctx* ctx, const void* proto, size_t len) {
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
    ctx->uni_fast = ak_uni_wordwise(h);
    if (ctx->uni_fast) {
        const std::vector<unsigned long long> img = ak_build_uni_image(h, AKC_BITS);
        if ((rc = ak_install_cache(ctx, ctx->uni_allocs, img, ctx->tok_cache[1]))) return rc;
    }
    for (int form = 0; form < 2; ++form)
        if ((rc = ak_install_decode(ctx, ctx->uni_allocs, ak_build_spm_decode(h, form), ctx->dec[1][form]))) return rc;
    ctx->uni_d = d;
    ctx->uni_h = std::move(h);
    ctx->has_uni = true;
    return AKSHAR_OK;
}

extern "C" {

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

// ---- event-stream encoders (ak_tok_kernels.cuh): words -> (row fix) -> lookup -----------------------------------------
struct AkTokOut {
    void* ids;
    int64_t id_cap;
    int ids_u16;
    void* splits;
    int splits_i32;
};
static int ak_run_tok(akshar_ctx* ctx, AkCall& C, const AkBatch& B, int64_t max_bytes, int kind, const AkTokOut& O) {
    int rc;
    const AkTokWs W = ak_tok_ws(max_bytes, B.n_rows);
    char* base = C.ws + C.L.scratch;
    const size_t avail = C.ws_bytes - C.L.scratch;
    // A workspace larger than the minimum: a quarter of the surplus each enlarges the id / scratch pool (uncached words,
    // exact Viterbi scratch) and the long-word pool (AKSHAR_ST_WORD asks for them), the rest gives every warp tile more
    // event slots (AKSHAR_ST_OVERFLOW with result[3] = the slots a warp tile needed)
    const size_t surplus = avail > W.total ? ((avail - W.total) / 4) & ~(size_t)255 : 0;
    const size_t pool_ints = W.pool_ints + surplus / 4, longpool_ints = W.longpool_ints + surplus / 4;
    char* longpool = base + W.longpool + surplus;
    char* sl = base + W.slots + 2 * surplus;
    const size_t left = avail - (W.slots + 2 * surplus) - 4096;
    // slots per warp tile: the largest power of two the workspace holds
    const double per_wt = (double)left / ((double)W.n_wt * (AKT_SLOT_BYTES + 0.001));
    int cap = AKT_CAP_MIN, shift = 8;
    while (cap < AKT_CAP_MAX && (double)(2 * cap) <= per_wt) { cap *= 2; ++shift; }
    const size_t n_slots = (size_t)W.n_wt * (size_t)cap;
    AkSlots S;
    S.ev = (AkEvent*)sl;
    S.count = (uint32_t*)(base + W.count);
    S.cap = cap;
    S.shift = shift;
    unsigned long long* resolved = (unsigned long long*)(sl + ak_align(n_slots * 8));
    uint32_t* aux = (uint32_t*)(sl + 2 * ak_align(n_slots * 8));
    int32_t* wt_ids = (int32_t*)(base + W.wt_ids);
    int64_t* wt_base = (int64_t*)(base + W.wt_base);
    unsigned long long* wt_seg = (unsigned long long*)(base + W.wt_seg);
    float* wt_segx = (float*)(base + W.wt_segx);
    const size_t nscan = ak_align(((size_t)W.n_wt / AKS_TILE + 2) * 8);
    unsigned long long* scan_state0 = (unsigned long long*)(base + W.scan_state);
    unsigned long long* scan_state1 = (unsigned long long*)(base + W.scan_state + nscan);
    AkcTable& tc = ctx->tok_cache[kind];
    if (!ctx->wc_hold || tc.reset) {
        // a quarter of the table in learned words is where probe chains start to grow
        const unsigned long long limit = (1ull << tc.work.bits) / 4;
        ak_cache_guard_kernel<<<ctx->sm_count * 4, 256, 0, C.stream>>>(tc.work.e, tc.image, tc.bytes / 8, tc.work.inserted, limit, tc.reset ? 1 : 0);
        if ((rc = ak_after_launch(ctx, "tok-cache-guard"))) return rc;
        ak_cache_guard_reset_kernel<<<1, 1, 0, C.stream>>>(tc.work.inserted, limit, tc.reset ? 1 : 0);
        if ((rc = ak_after_launch(ctx, "tok-cache-guard-reset"))) return rc;
        tc.reset = false;
    }
    AK_CUDA(ctx, cudaMemsetAsync(base + W.row_flag, 0, ak_align((size_t)B.n_rows + 1), C.stream));
    AK_CUDA(ctx, cudaMemsetAsync(scan_state0, 0, 2 * nscan, C.stream));
    int* tickets = (int*)C.ws;
    unsigned int* any_flag = (unsigned int*)(C.ws + 144);
    unsigned long long* pool_used = (unsigned long long*)(C.ws + 152);
    const int64_t base0 = B.text_begin - (int64_t)(((uintptr_t)B.text + (uintptr_t)B.text_begin) & 15u);
    const int64_t span = (B.dyn_end ? max_bytes : B.text_end) - base0;
    const int64_t n_wt_ub = (span + AKT_WARP_BYTES) / AKT_WARP_BYTES;
    AkWordsArgs A;
    A.B = B;
    A.T = ctx->T;
    A.bpe = ctx->bpe_d;
    A.base0 = base0;
    A.wrow = (const int64_t*)(base + W.wrow);
    A.S = S;
    A.row_flag = (uint8_t*)(base + W.row_flag);
    A.any_flag = any_flag;
    A.row_ev = (uint32_t*)(base + W.row_ev);
    long long* n_wt_dev = (long long*)(C.ws + 136);
    A.n_wt_out = n_wt_dev;
    const int entries = (int)(n_wt_ub * 2 + 3);
    ak_warp_rows_kernel<<<(entries + 255) / 256, 256, 0, C.stream>>>(B, base0, entries, (int64_t*)(base + W.wrow));
    if ((rc = ak_after_launch(ctx, "tok-warp-rows"))) return rc;
    if (kind == 1) {
        ak_long_rows_kernel<<<ctx->sm_count * 4, 256, 0, C.stream>>>(B, AKT_LONG_ROW, A.row_flag, any_flag);
        if ((rc = ak_after_launch(ctx, "tok-long-rows"))) return rc;
    }
    {
        AkTimed tm(ctx, AKSHAR_TIMER_WORDS, C.stream);
        const int g = ak_grid(ctx, ctx->occ_words[kind], (int)((n_wt_ub + AKW_THREADS / 32 - 1) / (AKW_THREADS / 32)));
        if (kind == 0) ak_words_kernel<0><<<g, AKW_THREADS, 0, C.stream>>>(A);
        else ak_words_kernel<1><<<g, AKW_THREADS, 0, C.stream>>>(A);
    }
    if ((rc = ak_after_launch(ctx, "tok-words"))) return rc;
    AkLookupCtx X;
    X.M.kind = kind;
    X.M.bpe = ctx->bpe_d;
    X.M.uni = ctx->uni_d;
    X.M.cache = tc.work;
    X.M.pool.base = (int32_t*)longpool;
    X.M.pool.used = (unsigned long long*)(C.ws + 128);
    X.M.pool.cap = longpool_ints;
    X.M.T = ctx->T;
    X.text = nullptr;                  // text / offsets / row count / result come from the (resolved) batch in the kernels
    X.off = nullptr;
    X.n_rows = 0;
    X.tb = X.te = 0;
    X.result = nullptr;
    X.ids = O.ids;
    X.id_cap = O.id_cap;
    X.ids_u16 = O.ids_u16;
    X.splits = O.splits;
    X.splits_i32 = O.splits_i32;
    X.row_flag = A.row_flag;
    X.row_fix = (unsigned long long*)(base + W.row_fix);
    X.pool = (int32_t*)(base + W.pool);
    X.pool_used = pool_used;
    X.pool_cap = pool_ints;
    X.any_fix = 0;
    AkRowFixArgs R;
    R.B = B;
    R.X.M = X.M;
    R.X.text = nullptr;
    R.X.off = nullptr;
    R.X.n_rows = 0;
    R.X.result = nullptr;
    R.X.ev = S.ev;
    R.X.n_events = 0;
    R.X.row_ev = A.row_ev;
    R.X.row_fix = (unsigned long long*)(base + W.row_fix);
    R.X.pool = X.pool;
    R.X.pool_used = pool_used;
    R.X.pool_cap = pool_ints;
    R.base0 = base0;
    R.cap = cap;
    R.row_flag = A.row_flag;
    R.any_flag = any_flag;
    ak_rowfix_kernel<<<ctx->sm_count * 8, 128, 0, C.stream>>>(R);
    if ((rc = ak_after_launch(ctx, "tok-rowfix"))) return rc;
    const int warp_ctas = (int)((n_wt_ub + AKL_THREADS / 32 - 1) / (AKL_THREADS / 32));      // CTAs when every warp takes one warp tile
    const int scan_tiles = (int)(n_wt_ub / AKS_TILE + 1);
    unsigned int* status_word = (unsigned int*)&B.result[2];
    AkResolveArgs Rs;
    Rs.B = B;
    Rs.X = X;
    Rs.base0 = base0;
    Rs.S = S;
    Rs.resolved = resolved;
    Rs.aux = aux;
    Rs.wt_ids = wt_ids;
    Rs.wt_seg = wt_seg;
    Rs.any_flag = any_flag;
    {
        AkTimed tm(ctx, kind == 0 ? AKSHAR_TIMER_BPE_ENCODE : AKSHAR_TIMER_UNIGRAM, C.stream);
        const int g = ak_grid(ctx, ctx->occ_resolve[kind], warp_ctas);
        if (kind == 0) ak_resolve_kernel<0><<<g, AKR_THREADS, 0, C.stream>>>(Rs);
        else ak_resolve_kernel<1><<<g, AKR_THREADS, 0, C.stream>>>(Rs);
    }
    if ((rc = ak_after_launch(ctx, "tok-resolve"))) return rc;
    if (kind == 1) {
        ak_scan_seg_kernel<<<ak_grid(ctx, 4, scan_tiles), AKS_THREADS, 0, C.stream>>>(wt_seg, n_wt_dev, wt_segx, tickets + 4, scan_state1, status_word);
        if ((rc = ak_after_launch(ctx, "tok-scan-seg"))) return rc;
        AkCheckArgs Ck;
        Ck.B = B;
        Ck.X = X;
        Ck.base0 = base0;
        Ck.S = S;
        Ck.resolved = resolved;
        Ck.aux = aux;
        Ck.wt_seg_before = wt_segx;
        Ck.wt_ids = wt_ids;
        Ck.any_flag = any_flag;
        ak_unicheck_kernel<<<ak_grid(ctx, ctx->occ_check, warp_ctas), AKL_THREADS, 0, C.stream>>>(Ck);
        if ((rc = ak_after_launch(ctx, "tok-check"))) return rc;
    }
    ak_scan_counts_kernel<<<ak_grid(ctx, 4, scan_tiles), AKS_THREADS, 0, C.stream>>>(wt_ids, 0, n_wt_dev, 1, wt_base, B.totals, tickets + 5, scan_state0,
                                                                                    status_word);
    if ((rc = ak_after_launch(ctx, "tok-scan-ids"))) return rc;
    AkEmitArgs E;
    E.B = B;
    E.X = X;
    E.base0 = base0;
    E.S = S;
    E.resolved = resolved;
    E.wt_base = wt_base;
    E.any_flag = any_flag;
    {
        AkTimed tm(ctx, AKSHAR_TIMER_EMIT, C.stream);
        if (O.ids_u16) ak_emit_kernel<uint16_t><<<ak_grid(ctx, ctx->occ_emit, warp_ctas), AKL_THREADS, 0, C.stream>>>(E);
        else ak_emit_kernel<int32_t><<<ak_grid(ctx, ctx->occ_emit, warp_ctas), AKL_THREADS, 0, C.stream>>>(E);
    }
    return ak_after_launch(ctx, "tok-emit");
}

// BPE over batch B (B may carry dyn_end from an earlier stage of a pipeline): the event-stream encoder in tile mode, the
// exact span walker (+ its conditional NFC passes) in row mode
// The BPE encoder has no bounded look-back and no row-sequential twin: both modes run the event-stream path, whose exact
// row kernel takes whatever the fast lanes hand over (added tokens, NFKC, rows that are not in NFC).
static int ak_run_bpe(akshar_ctx* ctx, AkCall& C, const AkBatch& B, int64_t max_bytes, const AkTokOut& O) {
    return ak_run_tok(ctx, C, B, max_bytes, 0, O);
}

static int ak_run_unigram(akshar_ctx* ctx, AkCall& C, const AkBatch& B, int64_t max_bytes, const AkTokOut& O, int mode) {
    if (mode == AKSHAR_MODE_TILES && ctx->uni_fast) return ak_run_tok(ctx, C, B, max_bytes, 1, O);
    if (O.ids_u16 || O.splits_i32) {
        ctx->err = "compact outputs need AKSHAR_MODE_TILES and a model of the shape scripts/train_spm.py writes";
        return AKSHAR_E_ARG;
    }
    int32_t* d_ids = (int32_t*)O.ids;
    int64_t* d_id_splits = (int64_t*)O.splits;
    const int64_t id_capacity = O.id_cap;
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
    AkDeviceGuard device_guard(ctx ? ctx->device : 0);
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
    const AkTokOut O = {d_ids, id_capacity, 0, d_id_splits, 0};
    return ak_run_bpe(ctx, C, C.B, text_end - text_begin, O);
}

int akshar_encode_unigram_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                                int64_t text_begin, int64_t text_end, int mode, int32_t* d_ids, int64_t id_capacity,
                                int64_t* d_id_splits, int64_t* d_result, void* d_workspace, size_t workspace_bytes,
                                void* stream) {
    AkDeviceGuard device_guard(ctx ? ctx->device : 0);
    AkCall C;
    if (mode != AKSHAR_MODE_TILES && mode != AKSHAR_MODE_ROWS) return AKSHAR_E_ARG;
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
    const AkTokOut O = {d_ids, id_capacity, 0, d_id_splits, 0};
    return ak_run_unigram(ctx, C, C.B, text_end - text_begin, O, mode);
}

int akshar_tokenizer_encode_batch(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                                  int64_t text_begin, int64_t text_end, uint32_t norm_flags, int kind, int mode,
                                  uint8_t* d_norm_text, int64_t norm_capacity, int64_t* d_norm_row_offsets, int32_t* d_ids,
                                  int64_t id_capacity, int64_t* d_id_splits, int64_t* d_result, void* d_workspace,
                                  size_t workspace_bytes, void* stream) {
    return akshar_tokenizer_encode_batch_ex(ctx, d_text, d_row_offsets, n_rows, text_begin, text_end, norm_flags, kind, mode,
                                            d_norm_text, norm_capacity, d_norm_row_offsets, d_ids, id_capacity, d_id_splits, 0u,
                                            d_result, d_workspace, workspace_bytes, stream);
}

int akshar_tokenizer_encode_batch_ex(akshar_ctx* ctx, const uint8_t* d_text, const int64_t* d_row_offsets, int64_t n_rows,
                                     int64_t text_begin, int64_t text_end, uint32_t norm_flags, int kind, int mode,
                                     uint8_t* d_norm_text, int64_t norm_capacity, int64_t* d_norm_row_offsets, void* d_ids,
                                     int64_t id_capacity, void* d_id_splits, uint32_t out_flags, int64_t* d_result,
                                     void* d_workspace, size_t workspace_bytes, void* stream) {
    AkDeviceGuard device_guard(ctx ? ctx->device : 0);
    if (!ctx) return AKSHAR_E_ARG;
    if (norm_capacity < 0 || !d_norm_row_offsets || (!d_norm_text && norm_capacity > 0) || (norm_flags & ~15u) ||
        (kind != 0 && kind != 1) || !d_id_splits || id_capacity < 0 || (!d_ids && id_capacity > 0) || (out_flags & ~3u)) {
        ctx->err = "bad argument";
        return AKSHAR_E_ARG;
    }
    if ((out_flags & AKSHAR_OUT_IDS_U16) && akshar_vocab_size(ctx, kind) > 65536) {
        ctx->err = "uint16 ids need a vocabulary of at most 65536 entries";
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
    if (n_rows == 0) {
        // splits[0] = 0 in either width
        AK_CUDA(ctx, cudaMemsetAsync(d_id_splits, 0, sizeof(int64_t), C.stream));
        return ak_empty_rows(ctx, d_norm_row_offsets, nullptr, C.stream);
    }
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
    const AkTokOut O = {d_ids, id_capacity, (out_flags & AKSHAR_OUT_IDS_U16) ? 1 : 0, d_id_splits, (out_flags & AKSHAR_OUT_SPLITS_I32) ? 1 : 0};
    if (kind == 0) return ak_run_bpe(ctx, C, B2, max_bytes, O);
    return ak_run_unigram(ctx, C, B2, max_bytes, O, mode);
}

}  // extern "C"

