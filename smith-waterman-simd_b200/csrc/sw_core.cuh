// sw_core.cuh -- the per-thread Smith-Waterman recurrence of the B200 kernel.
//
// What it computes: the value of the reference's scalar `SmithWaterman`
// (/root/reference/source.cpp:35-60) -- and therefore of SmithWaterman_simd ..
// SmithWaterman_simd9 (source.cpp:62-1071) on their common domain -- for TWO
// independent 128x128 pairs at once, one per 16-bit half of every 32-bit register.
//
// This is not a translation of the AVX2 code.  What is kept from the reference is the
// idea its README.md:2 names ("parallelogram"): cells of one anti-diagonal are mutually
// independent.  What is different:
//   * SIMD-in-register is across PAIRS (low half = pair 2t, high half = pair 2t+1), not
//     across rows of one pair, so no lane shifts (the reference's alignr/PRMT) exist at all.
//   * One thread owns a 16-row strip and sweeps it along anti-diagonals; the 16 cells of a
//     step are 16 independent dependency chains (ILP), and the parallelogram's triangular
//     ends are filled by the NEXT strip of the same pair ("wrap"), so only 240 padding
//     cells per pair are ever computed (1.4 %; the reference's scheme computes 12.5 %).
//   * Fast path = anti-diagonal offset DP.  With H~(i,j) = H(i,j) + g*(i+j) the up and
//     left moves cost nothing and a cell is TWO instructions:
//         t  = max(diag + s'', up)            VIADDMNMX.S16x2      s'' = max(s,-2g) + 2g >= 0
//         H~ = max(t, left, Z)                VIMNMX3.S16x2        Z = g*(i+j) (true zero)
//     (a diagonal step worth less than two gaps can never win, so clamping s at -2g does
//     not change any H).  The reference's simd9 (source.cpp:985-995) offsets only
//     vertically and still pays a subtraction per cell.
//   * Substitution scores: one PRMT per word.  Row k keeps S[a[k]][0..3] for both pairs as
//     8 bytes (two registers); the column's selector picks one byte per pair and
//     sign-replicates it into the upper byte of each half.  This is the register analogue
//     of the reference's pshufb table (source.cpp:518).
//   * The strip's bottom row travels to the next strip through a per-thread FIFO in
//     shared memory (the reference's `yoko` buffer, source.cpp:495-496,553).
//
// Every function here is `SWB_HD`: the same text compiles for the device (real packed
// instructions) and for the host (plain C emulation of them).  The host build exists ONLY
// for tests/emu (algorithm validation without a GPU); the product never runs it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SWB_HD __host__ __device__ __forceinline__
#else
#define SWB_HD inline
#endif

namespace swb {

// ------------------------------------------------------------------ packed primitives
#if defined(__CUDA_ARCH__)
SWB_HD uint32_t vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }                          // VIADD.16x2
SWB_HD uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }                         // VIMNMX.S16x2
SWB_HD uint32_t vaddmax2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }       // VIADDMNMX.S16x2
SWB_HD uint32_t vaddmax2_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }
SWB_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }    // VIMNMX3.S16x2
SWB_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));   // generic mode: nibble msb = replicate sign
    return d;
}
#else
SWB_HD int16_t lo16(uint32_t x) { return (int16_t)(x & 0xffffu); }
SWB_HD int16_t hi16(uint32_t x) { return (int16_t)(x >> 16); }
SWB_HD uint32_t pk(int32_t lo, int32_t hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
SWB_HD int32_t mx(int32_t a, int32_t b) { return a > b ? a : b; }
SWB_HD int32_t w16(int32_t a) { return (int16_t)a; }   // wrap-around, like the hardware
SWB_HD uint32_t vadd2(uint32_t a, uint32_t b) { return pk(lo16(a) + lo16(b), hi16(a) + hi16(b)); }
SWB_HD uint32_t vmax2(uint32_t a, uint32_t b) { return pk(mx(lo16(a), lo16(b)), mx(hi16(a), hi16(b))); }
SWB_HD uint32_t vaddmax2(uint32_t a, uint32_t b, uint32_t c)
{
    return pk(mx(w16(lo16(a) + lo16(b)), lo16(c)), mx(w16(hi16(a) + hi16(b)), hi16(c)));
}
SWB_HD uint32_t vaddmax2_relu(uint32_t a, uint32_t b, uint32_t c)
{
    return pk(mx(mx(w16(lo16(a) + lo16(b)), lo16(c)), 0), mx(mx(w16(hi16(a) + hi16(b)), hi16(c)), 0));
}
SWB_HD uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmax2(a, b), c); }
SWB_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t nib = (s >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(src >> (8 * (nib & 7))) & 0xff;
        if (nib & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        d |= byte << (8 * i);
    }
    return d;
}
#endif

// ------------------------------------------------------------------ kernel constants
// Filled on the host by sw_make_params() (sw_params.h) from the 4x4 matrix and the gap.
struct SwParams {
    uint32_t t4[4];    // t4[a] = bytes e(a,0..3): fast path e = max(S,-2g)+2g (0..127); general path e = S
    uint32_t dummy;    // profile word of an inert row (its cells never exceed a real neighbour)
    uint32_t G;        // fast: (g,g)            general: unused
    uint32_t NG;       // general: (-g,-g)       fast: unused
    uint32_t N2G;      // fast: (-2g,-2g)
    uint32_t K;        // fast: (112g,112g) FIFO re-basing; general: 0
    int32_t  gap;      // g
    int32_t  fast;     // 1 = anti-diagonal offset DP is exact for this matrix/gap
};

constexpr int SW_L = 128;          // sequence length (source.cpp:36-37)
constexpr int SW_R = 16;           // strip height = cells per anti-diagonal step per thread
constexpr int SW_STRIPS = SW_L / SW_R;
constexpr int SW_ITERS = SW_STRIPS * (SW_L / 16) + 1;   // 16-step iterations incl. the draining one

// Thread-private state.  Everything is indexed by compile-time constants after unrolling,
// so it lives in registers.
struct SwState {
    uint32_t h1[SW_R];    // row k at step T-1  (left of the cell being computed; up of row k+1)
    uint32_t h2[SW_R];    // row k at step T-2  (diag of row k+1)
    uint32_t prA[SW_R];   // S[a_lo[k]][0..3] as 4 bytes   (pair in the low half)
    uint32_t prB[SW_R];   // S[a_hi[k]][0..3] as 4 bytes   (pair in the high half)
    uint32_t sel[SW_R];   // PRMT selectors of the 16 most recent columns
    uint32_t up0, dg0;    // row 0's up / diag (from the FIFO)
    uint32_t Z;           // fast: packed g*(T+2) = the value of a true zero at step T
    uint32_t B;           // running best (fast: in the frame of the current step)
};

// byte `idx` (0..3) of word w, times 4 (a word offset into t4[]), masked to a valid code
SWB_HD uint32_t code_x4(uint32_t w, int idx)
{
    return (idx == 0) ? ((w << 2) & 0xcu) : ((w >> (8 * idx - 2)) & 0xcu);
}

// One 16-step iteration.  WRAP = this iteration starts at a step that is a multiple of 128:
// at sub-step u, row u leaves its strip and enters column 0 of the next one.
//   Fifo: pop(c) / push(c, v) with c the column 0..127.
//   bw_lo/bw_hi: the 16 target bases (bytes) of this iteration's columns, 4 words each.
//   aw_lo/aw_hi: (WRAP only) the 16 query bases of the strip being entered.
template <bool FAST, bool WRAP, class Fifo, class Table>
SWB_HD void sw_iter16(SwState& st, Fifo& fifo, const Table& t4, const SwParams& prm, int col0,
                      const uint32_t (&bw_lo)[4], const uint32_t (&bw_hi)[4],
                      const uint32_t (&aw_lo)[4], const uint32_t (&aw_hi)[4], bool next_is_dummy)
{
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        // --- selector of the column entering the window: bytes (b_lo, b_hi) -> nibbles
        //     (b_lo, 8|b_lo, 4|b_hi, 12|b_hi): low half <- byte b_lo of prA sign-extended,
        //     high half <- byte b_hi of prB sign-extended.
        {
            const uint32_t pick = 0x4440u | (uint32_t)(u & 3) | ((uint32_t)(u & 3) << 4);
            const uint32_t w = prmt(bw_lo[u >> 2], bw_hi[u >> 2], pick);   // bytes 2,3 are don't-care: PRMT reads selector bits 0..15 only
            st.sel[u] = w * 17u + 0xC480u;                                  // IMAD: the FMA pipe, not the ALU pipe
        }
        if (WRAP) {   // row u enters the next strip: new query profile
            if (next_is_dummy) {
                st.prA[u] = prm.dummy;
                st.prB[u] = prm.dummy;
            } else {
                st.prA[u] = t4(code_x4(aw_lo[u >> 2], u & 3));
                st.prB[u] = t4(code_x4(aw_hi[u >> 2], u & 3));
            }
        }
        // --- row 0's upper neighbours come from the previous strip's bottom row
        const uint32_t popped = fifo.pop(WRAP ? u : col0 + u);
        st.dg0 = st.up0;
        st.up0 = FAST ? vadd2(popped, prm.K) : popped;
        const uint32_t Zm2 = FAST ? vadd2(st.Z, prm.N2G) : 0u;   // true zero two steps ago (WRAP only)

        uint32_t hn[SW_R];
#pragma unroll
        for (int k = 0; k < SW_R; ++k) {
            const uint32_t s = prmt(st.prA[k], st.prB[k], st.sel[(u - k) & 15]);
            uint32_t up = (k == 0) ? st.up0 : st.h1[k - 1];
            uint32_t dg = (k == 0) ? st.dg0 : st.h2[k - 1];
            uint32_t lf = st.h1[k];
            if (WRAP && k == u) {   // column 0 of a strip: H[i][-1] = H[i-1][-1] = 0  (source.cpp:44 zero-initialised table)
                lf = 0u;
                dg = FAST ? Zm2 : 0u;
            }
            if (FAST) {
                const uint32_t t = vaddmax2(dg, s, up);
                hn[k] = vmax3(t, lf, st.Z);
            } else {
                const uint32_t t = vadd2(vmax2(up, lf), prm.NG);
                hn[k] = vaddmax2_relu(dg, s, t);
            }
        }
        // --- bottom row to the FIFO (column of row 15 at this step)
        //     (a WRAP iteration always starts at column 0; elsewhere col0 >= 16 and nothing wraps,
        //     so every FIFO address below is `per-iteration base + compile-time offset`)
        fifo.push(WRAP ? ((u - (SW_R - 1)) & (SW_L - 1)) : (col0 + u - (SW_R - 1)), hn[SW_R - 1]);
        // --- running best
        if (FAST) st.B = vaddmax2(st.B, prm.G, hn[0]);
        else      st.B = vmax2(st.B, hn[0]);
#pragma unroll
        for (int k = 1; k + 1 < SW_R; k += 2) st.B = vmax3(st.B, hn[k], hn[k + 1]);
        st.B = vmax2(st.B, hn[SW_R - 1]);
        // --- advance
#pragma unroll
        for (int k = 0; k < SW_R; ++k) { st.h2[k] = st.h1[k]; st.h1[k] = hn[k]; }
        if (FAST) st.Z = vadd2(st.Z, prm.G);
    }
}

// 8 bytes from a byte pointer that is 8-byte aligned
SWB_HD void ld8(const uint8_t* p, uint32_t& w0, uint32_t& w1)
{
#if defined(__CUDA_ARCH__)
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    w0 = v.x; w1 = v.y;
#else
    w0 = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    w1 = (uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
#endif
}

SWB_HD void ld16(const uint8_t* p, uint32_t (&w)[4])
{
    ld8(p, w[0], w[1]);
    ld8(p + 8, w[2], w[3]);
}

// Scores two pairs: (a_lo,b_lo) in the low halves, (a_hi,b_hi) in the high halves.
// All four pointers address 128 byte-coded bases (0..3), 8-byte aligned.
// The FIFO must hold 128 words for this thread; it is (re)initialised here.
template <bool FAST, class Fifo, class Table>
SWB_HD void sw128_two_pairs(const uint8_t* a_lo, const uint8_t* a_hi, const uint8_t* b_lo, const uint8_t* b_hi,
                            Fifo& fifo, const Table& t4, const SwParams& prm, int32_t& score_lo, int32_t& score_hi)
{
    SwState st;
#pragma unroll
    for (int k = 0; k < SW_R; ++k) {
        st.h1[k] = 0u; st.h2[k] = 0u; st.sel[k] = 0u;
        st.prA[k] = prm.dummy; st.prB[k] = prm.dummy;
    }
    // Top boundary H[0][*] = 0 (source.cpp:44).  Fast path: the word popped for column c at
    // step T=c must read g*(c+1) after re-basing by K = 112g, i.e. g*(c-111).
    {
        const uint32_t step = FAST ? prm.G : 0u;
        uint32_t v = 0u;
        if (FAST) { const int32_t f = -111 * prm.gap; v = ((uint32_t)f & 0xffffu) * 0x10001u; }
        for (int c = 0; c < SW_L; ++c) { fifo.push(c, v); v = vadd2(v, step); }
    }
    st.Z = FAST ? vadd2(prm.G, prm.G) : 0u;   // g*(0+2)
    st.B = FAST ? prm.G : 0u;                 // best = 0 in the frame of step -1
    st.up0 = FAST ? prm.G : 0u;               // true zero at step -1 (only ever read as a diag of column 0, then overridden)
    st.dg0 = 0u;

    uint32_t bn_lo[4], bn_hi[4], an_lo[4], an_hi[4];
    ld16(b_lo, bn_lo); ld16(b_hi, bn_hi);
    ld16(a_lo, an_lo); ld16(a_hi, an_hi);

    for (int it = 0; it < SW_ITERS; ++it) {
        const int col0 = 16 * (it & 7);
        uint32_t bw_lo[4], bw_hi[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { bw_lo[q] = bn_lo[q]; bw_hi[q] = bn_hi[q]; }
        {   // prefetch the next iteration's target bases
            const int nc = 16 * ((it + 1) & 7);
            ld16(b_lo + nc, bn_lo); ld16(b_hi + nc, bn_hi);
        }
        if ((it & 7) == 0) {
            const int strip = it >> 3;
            uint32_t aw_lo[4], aw_hi[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { aw_lo[q] = an_lo[q]; aw_hi[q] = an_hi[q]; }
            {   // prefetch the following strip's query bases (clamped: the last two reads are unused)
                const int ns = (strip + 1 < SW_STRIPS) ? strip + 1 : SW_STRIPS - 1;
                ld16(a_lo + 16 * ns, an_lo); ld16(a_hi + 16 * ns, an_hi);
            }
            sw_iter16<FAST, true>(st, fifo, t4, prm, col0, bw_lo, bw_hi, aw_lo, aw_hi, strip >= SW_STRIPS);
        } else {
            sw_iter16<FAST, false>(st, fifo, t4, prm, col0, bw_lo, bw_hi, bw_lo, bw_hi, false);
        }
    }
    if (FAST) {
        // B is in the frame of the last step T = 16*SW_ITERS-1; st.Z is one step further.
        const int32_t zl = (int32_t)(int16_t)(st.Z & 0xffffu) - prm.gap;
        const int32_t zh = (int32_t)(int16_t)(st.Z >> 16) - prm.gap;
        score_lo = (int32_t)(int16_t)(st.B & 0xffffu) - zl;
        score_hi = (int32_t)(int16_t)(st.B >> 16) - zh;
    } else {
        score_lo = (int32_t)(int16_t)(st.B & 0xffffu);
        score_hi = (int32_t)(int16_t)(st.B >> 16);
    }
}

} // namespace swb
