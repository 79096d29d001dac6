// sw_params.h -- host-side derivation of the kernel constants from the reference's
// arguments `score_matrix` (4x4 int8, index seq1*4+seq2, source.cpp:38,50) and
// `gap_penalty` (int8, source.cpp:39).  Plain C++, no CUDA.
#pragma once
#include <stdint.h>
#include "sw_core.cuh"

namespace swb {

enum SwDomain { SW_DOMAIN_OK = 0, SW_DOMAIN_BAD_MATRIX = 1, SW_DOMAIN_BAD_GAP = 2 };

// The domain on which the reference's SIMD kernels equal its scalar one:
// every matrix entry in [-127,127] (an entry of -128 wraps in source.cpp:492) and gap in
// [0,127].  Outside it the reference itself is inconsistent, so the batch entry refuses.
inline int sw_check_domain(const int8_t sm[16], int gap)
{
    for (int i = 0; i < 16; ++i) if (sm[i] == -128) return SW_DOMAIN_BAD_MATRIX;
    if (gap < 0 || gap > 127) return SW_DOMAIN_BAD_GAP;
    return SW_DOMAIN_OK;
}

inline uint32_t sw_pack2(int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; }

// Whether sequences of length L can be scored at all in packed int16: the largest H is
// L * max(S) (source.cpp:47-53), which must fit.  True for every reference-domain matrix at
// L = 128 (128*127 = 16256) and L = 256; at L = 512 it needs max(S) <= 63.
inline bool sw_len_supported(const int8_t sm[16], int L)
{
    int smax = 0;
    for (int i = 0; i < 16; ++i) if (sm[i] > smax) smax = sm[i];
    return (L == 128 || L == 256 || L == 512) && L * smax <= 32767;
}

inline SwParams sw_make_params(const int8_t sm[16], int gap, int force_general = 0, int L = SW_L)
{
    SwParams p;
    int smax = 0;
    for (int i = 0; i < 16; ++i) if (sm[i] > smax) smax = sm[i];
    // Fast path (anti-diagonal offset DP) is exact iff every shifted score
    // e = max(S,-2g)+2g fits a non-negative signed byte (smax + 2g <= 127) and the largest
    // packed value L*smax + (L+2)*g fits int16.  The frame is renormalised every L steps; at
    // L = 128 the second condition follows from the first (128*(127-2g) + 130g <= 16256).
    const bool fast = !force_general && (smax + 2 * gap <= 127) && (L * smax + (L + 2) * gap <= 32767);
    p.fast = fast ? 1 : 0;
    p.gap = gap;
    for (int a = 0; a < 4; ++a) {
        uint32_t w = 0;
        for (int b = 0; b < 4; ++b) {
            int e = sm[a * 4 + b];
            if (fast) { if (e < -2 * gap) e = -2 * gap; e += 2 * gap; }
            w |= ((uint32_t)e & 0xffu) << (8 * b);
        }
        p.t4[a] = w;
    }
    p.dummy = fast ? 0u : 0x81818181u;   // fast: e = 0 (a step worth two gaps); general: S = -127
    p.G = sw_pack2(gap);
    p.NG = sw_pack2(-gap);
    p.C = sw_pack2(L * gap);
    p.one = 1u;
    p.mone = 0xffffffffu;
    return p;
}

} // namespace swb
