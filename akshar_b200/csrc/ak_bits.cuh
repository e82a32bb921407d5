// Parallel bit streams: the front end of the bit-parallel ("v3") kernels.
//
// A lane owns 32 consecutive text bytes (two 16-byte loads).  Instead of walking them byte by byte it transposes
// them into 8 BASIS PLANES -- P[k] bit i = bit k of byte i -- with three butterfly steps per 8 bytes, after which
//   * every byte class is a handful of LOP3s on whole planes (32 bytes per instruction, no table, no divergence),
//   * "the same byte as k positions earlier" is the AND over the planes of ~(P ^ (P << k)),
//   * code-point structure is mask arithmetic: a bit moves from one lead byte to the next through the run of
//     continuation bytes between them by a single integer add (carry propagation).
// Cross-lane context travels as a few packed bits per shuffle.  All helpers are AK_HD so that
// tests/csrc/host_harness.cpp runs the identical arithmetic lane by lane on the CPU.
#pragma once
#include "ak_unicode.cuh"

AK_HD uint32_t akb_prmt(uint32_t a, uint32_t b, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, s);
#else
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
#endif
}
// (hi << s) | (lo >> (32 - s)), 0 < s < 32
AK_HD uint32_t akb_fsl(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    return (hi << s) | (lo >> (32u - s));
#endif
}
// (lo >> s) | (hi << (32 - s)), 0 < s < 32
AK_HD uint32_t akb_fsr(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    return (lo >> s) | (hi << (32u - s));
#endif
}
AK_HD int akb_ctz(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
AK_HD int akb_clz(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}
AK_HD int akb_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

// 8x8 bit-matrix transpose of the 8 bytes (lo = bytes 0..3, hi = bytes 4..7): afterwards byte k holds bit k of
// every input byte (bit p of byte k = bit k of input byte p).  Index of (row r, column c) = 8 r + c; the first two
// butterflies never cross the 32-bit halves.
AK_HD void akb_transpose8(uint32_t& lo, uint32_t& hi) {
    uint32_t t;
    t = (lo ^ (lo >> 7)) & 0x00AA00AAu; lo ^= t ^ (t << 7);
    t = (hi ^ (hi >> 7)) & 0x00AA00AAu; hi ^= t ^ (t << 7);
    t = (lo ^ (lo >> 14)) & 0x0000CCCCu; lo ^= t ^ (t << 14);
    t = (hi ^ (hi >> 14)) & 0x0000CCCCu; hi ^= t ^ (t << 14);
    t = (lo ^ (hi << 4)) & 0xF0F0F0F0u;
    lo ^= t;
    hi ^= t >> 4;
}

// x[0..7]: the lane's 32 bytes (little-endian words) -> P[0..7]: the basis planes
AK_HD void akb_planes(const uint32_t* x, uint32_t* P) {
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        lo[g] = x[2 * g];
        hi[g] = x[2 * g + 1];
        akb_transpose8(lo[g], hi[g]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t sel = (uint32_t)k | ((uint32_t)(4 + k) << 4);
        P[k] = akb_prmt(akb_prmt(lo[0], lo[1], sel), akb_prmt(lo[2], lo[3], sel), 0x5410u);
        P[4 + k] = akb_prmt(akb_prmt(hi[0], hi[1], sel), akb_prmt(hi[2], hi[3], sel), 0x5410u);
    }
}

// M marks lead bytes; the result marks, for every marked lead, the NEXT lead (the bit rides the carry through the
// continuation bytes C between them).  cin = the mask's value at the last lead before the lane (0 / 1).
AK_HD uint32_t akb_fwd(uint32_t M, uint32_t C, uint32_t cin) {
    const uint32_t t = (M << 1) | cin;
    const uint32_t s = (t & C) + C;
    return (s | t) & ~C;
}
// the result marks, for every marked lead, the PREVIOUS lead.  cin = the mask's value at the first lead after the lane.
AK_HD uint32_t akb_bwd(uint32_t M, uint32_t C, uint32_t cin) {
    uint32_t t = (M >> 1) | (cin << 31);
    uint32_t r = t & ~C;
    t = (t & C) >> 1; r |= t & ~C;
    t = (t & C) >> 1; r |= t & ~C;
    t = (t & C) >> 1; r |= t & ~C;
    return r;
}
// value (0 / 1) of M at the highest / lowest set bit of L
AK_HD uint32_t akb_at_last(uint32_t M, uint32_t L) { return L ? ((M >> (31 - akb_clz(L))) & 1u) : 0u; }
AK_HD uint32_t akb_at_first(uint32_t M, uint32_t L) { return (M & (L & (0u - L))) ? 1u : 0u; }
