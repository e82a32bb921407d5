// Fast lane of normalize_text (reference normalize.py:117-148, default flags) for the hot kernel.
//
// Work unit: a 16-byte CHUNK of the concatenated text per thread (one coalesced 16-byte load), 32 chunks per warp of
// which lanes 1..30 are REAL (480 bytes) and lanes 0 / 31 are halo chunks that only provide context, so a warp needs
// nothing from other warps.  Ownership is the walker's (ak_text_core.cuh): a chunk owns the code points whose lead
// byte lies in it.  The fast lane handles a chunk when every code point it owns sits in an NFC segment that is
// not "troubled" (same definition as ak_scan_segment) and the little context it needs is available from its two
// neighbours; the result is then a 19-bit EMIT MASK over the chunk's bytes (+3 bytes of a straddling code point):
// the output is exactly the emitted source bytes, with A-Z lowered.  Everything else (a troubled segment, U+0130,
// context further away than one chunk) makes the lane SLOW: it runs the exact walker ak_norm_span on its 16 bytes.
// Both lanes implement the same per-code-point ownership + look-ahead collapse, so they compose bit-exactly.
//
// The three phases are AK_HD so that tests/csrc/host_harness.cpp can run the identical logic chunk by chunk.
#pragma once
#include "ak_text_core.cuh"

#define AKF_NONE 0xFFFFFFFFu
#define AKF_REAL 30
#define AKF_WARP_BYTES (AKF_REAL * 16)

// summary flags
#define AKF_TROUBLE 1u         // some owned code point is troubled / not handled by the fast lane
#define AKF_BOUNDARY 2u        // the chunk owns a head code point or contains a row start
#define AKF_LEAD_TROUBLE 4u    // a troubled code point before the first boundary (belongs to a segment headed earlier)
#define AKF_OPEN 8u            // neither a kept code point nor a row start: look-ahead cannot be answered here
#define AKF_ROWSTART 16u       // contains a row start
#define AKF_FIRST_DEP 32u      // first owned code point is a non-head that continues a segment from the previous chunk
#define AKF_KEPT_BEFORE_ROW 64u   // a kept code point precedes the first row start (so the incoming run state matters)

struct AkChunk {
    uint32_t w[5];        // the 16 bytes of the chunk + the 4 bytes that follow (little endian words)
    uint32_t rows;        // bit i: a row starts at byte i (16 bits)
    uint32_t own;         // bit i: byte i lies inside the text
    // ---- filled by phase A
    uint32_t cpv[16];     // code point (A-Z lowered) whose lead byte is byte i, for kept leads
    uint32_t kept;        // bit i: byte i is the lead of an owned code point that survives lowercase + filter
    uint32_t lead;        // bit i: byte i is the lead of an owned code point
    uint32_t flags;
    uint32_t first_w;     // props of the first owned lead when it is not on a row start (else AKF_NONE)
    uint32_t last_w;      // props of the last owned lead (AKF_NONE if none)
    uint32_t F;           // first kept code point before any row start (AKF_NONE if a row start comes first / none)
    uint32_t L1, L2;      // last and second-last kept code points since the last row start (AKF_NONE if absent)
};

AK_HD uint32_t akf_byte(const AkChunk& c, int i) { return (c.w[i >> 2] >> ((i & 3) * 8)) & 0xFFu; }

// One shared-memory load for ASCII and U+0900-09FF (the closed alphabet of the hot path) through a single index
// computation -- no divergence between the two -- and the 2-stage global table for everything else.
AK_HD uint32_t akf_props(const AkTables& T, const uint32_t* lut, uint32_t cp) {
    const uint32_t d = cp - 0x900u;
    const uint32_t idx = cp < 0x80u ? cp : 128u + d;
    if (cp < 0x80u || d < 0x100u) return lut[idx];
    return ak_props(T, cp);
}

// refined per-code-point trouble test, identical to the loop body of ak_scan_segment; prev = props of the
// preceding code point of the same row
AK_HD bool akf_trouble_after(uint32_t prev, uint32_t w) {
    uint32_t qc = AK_QC(w), cc = AK_CCC(w);
    return qc == 1u || (qc == 2u && !AK_INERT_BASE(prev)) || (cc != 0u && AK_CCC(prev) > cc);
}

// ---- phase A: decode, classify, summarise (no context from other chunks) ------------------------------------
AK_HD void akf_phase_a(const AkTables& T, const uint32_t* lut, AkChunk& c) {
    uint32_t kept = 0, lead = 0, flags = 0;
    uint32_t prev_w = AKF_NONE;          // props of the previous owned code point of the same row in this chunk
    bool have_prev = false;              // prev_w valid (same row, this chunk)
    bool first_seen = false;
    bool boundary_seen = false;
    bool row_seen = false;
    uint32_t F = AKF_NONE, L1 = AKF_NONE, L2 = AKF_NONE, first_w = AKF_NONE, last_w = AKF_NONE;
    bool f_done = false;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        c.cpv[i] = 0;
        const bool row_here = (c.rows >> i) & 1u;
        if (row_here) {
            have_prev = false;
            boundary_seen = true;
            row_seen = true;
            f_done = true;               // look-ahead from the previous chunk stops at a row start
            L1 = L2 = AKF_NONE;
            flags |= AKF_ROWSTART | AKF_BOUNDARY;
        }
        const uint32_t b = akf_byte(c, i);
        if (!((c.own >> i) & 1u) || (b & 0xC0u) == 0x80u) continue;
        // decode (valid UTF-8 assumed; the 3 bytes after the chunk are in w[4])
        const uint32_t b1 = akf_byte(c, i + 1) & 0x3Fu, b2 = akf_byte(c, i + 2) & 0x3Fu, b3 = akf_byte(c, i + 3) & 0x3Fu;
        uint32_t cp;
        if (b < 0x80u) cp = b;
        else if (b < 0xE0u) cp = ((b & 0x1Fu) << 6) | b1;
        else if (b < 0xF0u) cp = ((b & 0x0Fu) << 12) | (b1 << 6) | b2;
        else cp = ((b & 0x07u) << 18) | (b1 << 12) | (b2 << 6) | b3;
        const uint32_t w = akf_props(T, lut, cp);
        lead |= 1u << i;
        last_w = w;
        const bool head = AK_NFC_HEAD(w);
        bool trouble;
        if (!first_seen && !row_here) {
            // continues the previous chunk's row: the context-dependent part of the test is resolved in phase B
            first_w = w;
            trouble = AK_QC(w) == 1u;
            if (!head) flags |= AKF_FIRST_DEP;
        } else if (!have_prev) {
            trouble = AK_QC(w) != 0u;                          // first code point of a row (ak_scan_segment's start)
        } else {
            trouble = akf_trouble_after(prev_w, w);
        }
        first_seen = true;
        if (cp == 0x130u) trouble = true;                      // lower() yields two code points: walker only
        if (trouble) {
            flags |= AKF_TROUBLE;
            if (!boundary_seen) flags |= AKF_LEAD_TROUBLE;
        }
        if (head) { boundary_seen = true; flags |= AKF_BOUNDARY; }
        prev_w = w;
        have_prev = true;
        if (AK_ALLOW(w)) {
            const uint32_t y = (cp - 'A' < 26u) ? cp + 32u : cp;
            c.cpv[i] = y;
            kept |= 1u << i;
            if (!f_done) { F = y; f_done = true; }
            if (!row_seen) flags |= AKF_KEPT_BEFORE_ROW;
            L2 = L1;
            L1 = y;
        }
    }
    if (!f_done) flags |= AKF_OPEN;
    c.kept = kept;
    c.lead = lead;
    c.flags = flags;
    c.first_w = first_w;
    c.last_w = last_w;
    c.F = F;
    c.L1 = L1;
    c.L2 = L2;
}

// ---- phase B1: the deferred test of the first owned code point, given the previous chunk's last props -------
AK_HD void akf_resolve_first(AkChunk& c, uint32_t prev_last_w) {
    if (c.first_w == AKF_NONE) return;
    bool t;
    if (prev_last_w == AKF_NONE) t = AK_QC(c.first_w) != 0u;          // nothing before it inside the text: row start
    else t = akf_trouble_after(prev_last_w, c.first_w);
    if (t) {
        c.flags |= AKF_TROUBLE;
        if (c.flags & AKF_FIRST_DEP) c.flags |= AKF_LEAD_TROUBLE;    // it is a non-head: part of the earlier segment
    }
}

// ---- phase B2: slow-lane decision from the two neighbours' (resolved) summaries ----------------------------
struct AkNeighbor {
    uint32_t flags, F, L1, L2;
};
AK_HD bool akf_is_slow(const AkChunk& c, const AkNeighbor& prev, const AkNeighbor& next) {
    if (c.flags & AKF_TROUBLE) return true;
    // leading partial segment headed in an earlier chunk
    if ((c.flags & AKF_FIRST_DEP) && ((prev.flags & AKF_TROUBLE) || !(prev.flags & AKF_BOUNDARY))) return true;
    // trailing segment continuing into the next chunk
    if ((next.flags & AKF_LEAD_TROUBLE) || !(next.flags & AKF_BOUNDARY)) return true;
    // incoming run state: the previous chunk must know the last two kept code points of the row
    if ((c.flags & AKF_KEPT_BEFORE_ROW) &&
        ((prev.flags & AKF_TROUBLE) || (!(prev.flags & AKF_ROWSTART) && prev.L2 == AKF_NONE)))
        return true;
    return false;
}

// ---- phase B3: elongation collapse -> emit mask (bits 0..18 over the chunk bytes + 3 straddling bytes) ------
// returns false when the held-back element at the chunk end cannot be resolved from `next` (-> slow lane)
AK_HD bool akf_collapse(const AkChunk& c, const AkNeighbor& prev, const AkNeighbor& next, uint32_t& emit_out) {
    uint32_t last = prev.L1;
    int n = (last == AKF_NONE) ? 0 : ((prev.L2 == last) ? 2 : 1);
    uint32_t emit = 0, pend = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if ((c.rows >> i) & 1u) {
            emit |= pend;
            pend = 0;
            last = AKF_NONE;
            n = 0;
        }
        if (!((c.kept >> i) & 1u)) continue;
        const uint32_t b = akf_byte(c, i);
        const uint32_t len = b < 0x80u ? 1u : b < 0xE0u ? 2u : b < 0xF0u ? 3u : 4u;
        const uint32_t bits = ((1u << len) - 1u) << i;
        const uint32_t y = c.cpv[i];
        if (y == last && y != 0x0Au) {
            if (n == 1) { n = 2; pend = bits; }
            else { n = 3; pend = 0; }
        } else {
            emit |= pend | bits;
            pend = 0;
            last = y;
            n = 1;
        }
    }
    if (pend) {
        if ((next.flags & AKF_OPEN) || ((next.flags & AKF_TROUBLE) && next.F != AKF_NONE)) return false;
        if (next.F != last) emit |= pend;
    }
    emit_out = emit;
    return true;
}

// ---- phase C: the emitted bytes, in order, to `out`; returns the count --------------------------------------
AK_HD int akf_write(const AkChunk& c, uint32_t emit, uint8_t* out) {
    int n = 0;
#pragma unroll
    for (int i = 0; i < 19; ++i) {
        if ((emit >> i) & 1u) {
            uint32_t b = akf_byte(c, i);
            if (b - 'A' < 26u) b += 32u;
            out[n++] = (uint8_t)b;
        }
    }
    return n;
}
