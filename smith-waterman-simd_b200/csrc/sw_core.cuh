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
    uint32_t G;        // fast: (g,g)
    uint32_t NG;       // general: (-g,-g)
    uint32_t C;        // fast: (L*g,L*g), subtracted from every live value once per strip (frame renormalisation)
    uint32_t one;      // 1          } multipliers the compiler cannot fold: `a*one + b` stays an IMAD, i.e. the
    uint32_t mone;     // 0xffffffff } frame bookkeeping adds run on the FMA pipe while the ALU pipe is saturated
    int32_t  gap;      // g
    int32_t  fast;     // 1 = anti-diagonal offset DP is exact for this matrix/gap
};

// a + b and a - b for frame bookkeeping: exact as plain 32-bit operations because both halves
// of every operand are non-negative and no half can carry or borrow (see SwState).
SWB_HD uint32_t fadd(uint32_t a, uint32_t b, const SwParams& p) { return a * p.one + b; }
SWB_HD uint32_t fsub(uint32_t a, uint32_t b, const SwParams& p) { return b * p.mone + a; }

constexpr int SW_L = 128;          // the reference's sequence length (source.cpp:36-37); the kernels are templated
                                   // on L (a multiple of 16) for the length sweep, BASELINE.json configs[3]
constexpr int SW_R = 16;           // strip height = cells per anti-diagonal step per thread

// Thread-private state.  Everything is indexed by compile-time constants after unrolling,
// so it lives in registers.
//
// Fast-path frame: at step T (counted from the last renormalisation) a register holds
// H + Z_T with Z_T = g*(T+2) in both halves; every value is >= 0, so the frame bookkeeping
// (Z += G, FIFO re-basing, renormalisation) is done with PLAIN 32-bit adds -- no carry can
// cross the halves -- which the compiler is free to issue on the FMA pipe (IMAD.IADD)
// instead of the saturated ALU pipe.  The FIFO itself holds frame-free true H values.
struct SwState {
    uint32_t h1[SW_R];    // row k at step T-1  (left of the cell being computed; up of row k+1)
    uint32_t h2[SW_R];    // row k at step T-2  (diag of row k+1)
    uint32_t prA[SW_R];   // S[a_lo[k]][0..3] as 4 bytes   (pair in the low half)
    uint32_t prB[SW_R];   // S[a_hi[k]][0..3] as 4 bytes   (pair in the high half)
    uint32_t sel[SW_R];   // PRMT selectors of the 16 most recent columns
    uint32_t up0, dg0;    // row 0's up / diag (from the FIFO)
    uint32_t Z, Zp;       // fast: Z_T and Z_{T-1}
    uint32_t B;           // running best (fast: in the frame of the current step)
    uint32_t bq_lo[2], bq_hi[2];   // target bases of the next iteration's columns 0..7 (loaded 8 steps ahead)
    uint32_t pq[16];               // Fifo::kPrefetch only: FIFO words of the next iteration, loaded ahead of use
                                   // (kPrefetch 1: its first 8 words; kPrefetch 2: all 16)
};

// byte `idx` (0..3) of word w, times 4 (a word offset into t4[]), masked to a valid code
SWB_HD uint32_t code_x4(uint32_t w, int idx)
{
    return (idx == 0) ? ((w << 2) & 0xcu) : ((w >> (8 * idx - 2)) & 0xcu);
}

// 8 bytes from a byte pointer that is 8-byte aligned.  COHERENT = false: the read-only path (ld.global.nc), for
// arrays nobody writes while the kernel runs.  COHERENT = true: a plain load -- the persistent consumer kernel
// (sw_feed_kernel.cuh) reads sequences that its own block expanded from the 2-bit wire format a moment ago.
template <bool COHERENT = false>
SWB_HD void ld8(const uint8_t* p, uint32_t& w0, uint32_t& w1)
{
#if defined(__CUDA_ARCH__)
    const uint2 v = COHERENT ? *reinterpret_cast<const uint2*>(p) : __ldg(reinterpret_cast<const uint2*>(p));
    w0 = v.x; w1 = v.y;
#else
    w0 = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    w1 = (uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
#endif
}

template <bool COHERENT = false>
SWB_HD void ld16(const uint8_t* p, uint32_t (&w)[4])
{
    ld8<COHERENT>(p, w[0], w[1]);
    ld8<COHERENT>(p + 8, w[2], w[3]);
}

// One 16-step iteration.  WRAP = this iteration starts at a step that is a multiple of 128:
// at sub-step u, row u leaves its strip and enters column 0 of the next one.
//   Fifo: pop(c) / push(c, v) with c the column 0..127.
//   b_lo: the low pair's target; the high pair's is at b_lo + dq.  This iteration covers
//         columns col0..col0+15; columns col0+8.. are loaded here at u = 0, the next
//         iteration's first eight at u = 8 (always 8 steps ahead of their first use).
//   aw_lo/aw_hi: (WRAP only) the 16 query bases of the strip being entered.
// V = variant bits (tuning knobs, all bit-exact):
//   bit 0 (SW_V_BEST_FMA): running best -- the +g frame shift is a plain add on the FMA pipe, then
//          8 VIMNMX3 (instead of VIADDMNMX + 7 VIMNMX3 + VIMNMX).  Measured -2.2 % time.
// Measured and REJECTED (profiles/r01/kbench_v3_variants.jsonl, kbench_v4_lds_hybrid.jsonl,
// kbench_v5_dual_best.jsonl):
//   * two independent best accumulators (half the dependency chain, same count): 0.7 % slower;
//   * IMAD/IMAD.HI byte extraction for the column selector instead of one PRMT: 4.6 % slower;
//   * substitution words of 4-10 rows of the strip from a lane-replicated shared-memory table
//     (IMAD address add + LDS on the FMA/LSU pipes instead of PRMT on the ALU pipe): bit-exact,
//     0.8-3.5 % slower than the all-PRMT kernel in the same block shape.
//   * n of the 16 rows computing t = max(diag + s'', up) on the FMA pipe as c + relu((a + b) - c) in fp16x2
//     (HADD2, HFMA2.RELU with a negated operand, HADD2): on bit patterns 0..2047 fp16 is one linear ramp
//     (subnormals + first binade, value = pattern * 2^-24), so these ARE the integer operations as long as every
//     value stays below 2048 -- the frame renormalised every 16 steps, both benchmark settings qualify.
//     Bit-exact on the B200 (reference checksum), SASS as intended (57 - n ALU-pipe, 7 + 3n FMA-pipe
//     instructions per step, 168 registers), but SLOWER by about 1 % per row (n = 2..12: 1.87 .. 2.07 ms against
//     1.79): three issue slots for one ALU-pipe slot is a bad trade even with the FMA pipe at 6 %
//     (profiles/r01/kbench_v6_fp16_rows_on_fma_pipe.jsonl; the patch is kept beside it).
//   bit 1 (SW_V_FIFO_PREOFF, fast path only, EXPERIMENTAL -- built by tools/kbench.cu and tests/emu only, not measured
//          yet): the FIFO holds each value already in the frame it will be popped in, so the pop needs no add.  Column c
//          is popped at step T = c of its epoch, where Z_{T-1} = g(c+1); a value pushed at step T (frame g(T+2)) for
//          column T-15 therefore goes in as H~ - 16g, and one pushed inside a wrap iteration at sub-step u < 15
//          (column L-15+u, popped later in the SAME epoch) as H~ + (L-16)g.  One IMAD less per step.
//   bit 2 (SW_V_COHERENT_LD): sequence loads are plain (coherent) loads instead of ld.global.nc -- no change to
//          the arithmetic; used by the persistent consumer kernel, whose blocks write the bytes they then read.
constexpr int SW_V_BEST_FMA = 1;
constexpr int SW_V_FIFO_PREOFF = 2;
constexpr int SW_V_COHERENT_LD = 4;

template <bool FAST, bool WRAP, int L, int V, class Fifo, class Table>
SWB_HD void sw_iter16(SwState& st, Fifo& fifo, const Table& t4, const SwParams& prm, int col0,
                      const uint8_t* b_lo, uint32_t dq,
                      const uint32_t (&aw_lo)[4], const uint32_t (&aw_hi)[4], bool next_is_dummy)
{
    uint32_t bw_lo[4], bw_hi[4];
    bw_lo[0] = st.bq_lo[0]; bw_lo[1] = st.bq_lo[1]; bw_hi[0] = st.bq_hi[0]; bw_hi[1] = st.bq_hi[1];
    {
        const uint8_t* p = b_lo + (WRAP ? 0 : col0) + 8;
        ld8<(V & SW_V_COHERENT_LD) != 0>(p, bw_lo[2], bw_lo[3]);
        ld8<(V & SW_V_COHERENT_LD) != 0>(p + dq, bw_hi[2], bw_hi[3]);
    }
    // A FIFO with long latency (global memory) is read ahead of use, like the bases.
    //   kPrefetch 1: 8..15 steps ahead -- words 8..15 of this iteration here, words 0..7 of the next at u = 8;
    //   kPrefetch 2: 16..31 steps ahead -- all 16 words of the next iteration here.
    // Every word read was pushed at least L-16 steps before its use by this same thread, and the
    // columns pushed meanwhile (col0-15 .. col0+15) never coincide with the ones prefetched (L >= 64).
    uint32_t pv[16];
    if (Fifo::kPrefetch == 1) {
#pragma unroll
        for (int q = 0; q < 8; ++q) pv[q] = st.pq[q];
#pragma unroll
        for (int q = 8; q < 16; ++q) pv[q] = fifo.pop((WRAP ? 0 : col0) + q);
    }
    if (Fifo::kPrefetch == 2) {
        const int nc = ((WRAP ? 0 : col0) + 16) & (L - 1);
#pragma unroll
        for (int q = 0; q < 16; ++q) { pv[q] = st.pq[q]; st.pq[q] = fifo.pop(nc + q); }
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        if (u == 8) {
            const int nc = ((WRAP ? 0 : col0) + 16) & (L - 1);
            const uint8_t* p = b_lo + nc;
            ld8<(V & SW_V_COHERENT_LD) != 0>(p, st.bq_lo[0], st.bq_lo[1]);
            ld8<(V & SW_V_COHERENT_LD) != 0>(p + dq, st.bq_hi[0], st.bq_hi[1]);
            if (Fifo::kPrefetch == 1) {
#pragma unroll
                for (int q = 0; q < 8; ++q) st.pq[q] = fifo.pop(nc + q);
            }
        }
        // --- selector of the column entering the window: bytes (b_lo, b_hi) -> nibbles
        //     (b_lo, 8|b_lo, 4|b_hi, 12|b_hi): low half <- byte b_lo of prA sign-extended,
        //     high half <- byte b_hi of prB sign-extended.
        {
            // pick bytes (b_lo, b_hi, 0, 0): nibbles 8|j replicate the sign of a code byte, i.e. 0
            const uint32_t pick = 0x8800u | (uint32_t)(u & 3) | ((uint32_t)((u & 3) + 4) << 4);
            const uint32_t w = prmt(bw_lo[u >> 2], bw_hi[u >> 2], pick);   // = b_lo + 256*b_hi
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
        // --- row 0's upper neighbours come from the previous strip's bottom row (frame-free in the FIFO)
        const uint32_t popped = Fifo::kPrefetch ? pv[u] : fifo.pop(WRAP ? u : col0 + u);
        st.dg0 = st.up0;
        if (FAST && (V & SW_V_FIFO_PREOFF)) st.up0 = popped;   // pushed in this step's frame already
        else st.up0 = FAST ? fadd(popped, st.Zp, prm) : popped;   // plain add: both halves >= 0, no carry across
        const uint32_t Zm2 = (FAST && WRAP) ? fsub(st.Zp, prm.G, prm) : 0u;   // true zero two steps ago

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
        // --- bottom row to the FIFO (column of row 15 at this step), as a true H value.
        //     (a WRAP iteration always starts at column 0; elsewhere col0 >= 16 and nothing wraps,
        //     so every FIFO address is `per-iteration base + compile-time offset`)
        if (FAST && (V & SW_V_FIFO_PREOFF)) {
            // hn >= Z_T >= 17g wherever 16g is taken off (T >= 15 there), and hn + (L-16)g stays inside the
            // int16 bound of sw_make_params (L*smax + (L+2)g): plain 32-bit arithmetic, no carry or borrow across
            const uint32_t w = (WRAP && u < SW_R - 1) ? fadd(hn[SW_R - 1], prm.G * (uint32_t)(L - 16), prm)
                                                      : fsub(hn[SW_R - 1], prm.G * 16u, prm);
            fifo.push(WRAP ? ((u - (SW_R - 1)) & (L - 1)) : (col0 + u - (SW_R - 1)), w);
        } else
        fifo.push(WRAP ? ((u - (SW_R - 1)) & (L - 1)) : (col0 + u - (SW_R - 1)),
                  FAST ? fsub(hn[SW_R - 1], st.Z, prm) : hn[SW_R - 1]);   // plain subtract: hn >= Z in both halves
        // --- running best
        if (FAST && (V & SW_V_BEST_FMA)) {
            st.B = fadd(st.B, prm.G, prm);      // best >= 0 in both halves: no carry across
#pragma unroll
            for (int k = 0; k + 1 < SW_R; k += 2) st.B = vmax3(st.B, hn[k], hn[k + 1]);
        } else {
            if (FAST) st.B = vaddmax2(st.B, prm.G, hn[0]);
            else      st.B = vmax2(st.B, hn[0]);
#pragma unroll
            for (int k = 1; k + 1 < SW_R; k += 2) st.B = vmax3(st.B, hn[k], hn[k + 1]);
            st.B = vmax2(st.B, hn[SW_R - 1]);
        }
        // --- advance
#pragma unroll
        for (int k = 0; k < SW_R; ++k) { st.h2[k] = st.h1[k]; st.h1[k] = hn[k]; }
        if (FAST) { st.Zp = st.Z; st.Z = fadd(st.Z, prm.G, prm); }
    }
}

// Scores two pairs: (a_lo,b_lo) in the low halves, (a_lo+dqa, b_lo+dqb) in the high halves
// (dqa = dqb = L for the neighbouring pair; 0 when the batch's last pair stands alone;
// dqb = 0 also when every pair shares one target, the one-vs-many entry).
// Both pointers address L byte-coded bases (0..3), 8-byte aligned; L is a power of two >= 32.
// The FIFO must hold L words for this thread; it is (re)initialised here.
template <bool FAST, int L, int V, class Fifo, class Table>
SWB_HD void sw_two_pairs(const uint8_t* a_lo, const uint8_t* b_lo, uint32_t dqa, uint32_t dqb,
                         Fifo& fifo, const Table& t4, const SwParams& prm, int32_t& score_lo, int32_t& score_hi)
{
    SwState st;
#pragma unroll
    for (int k = 0; k < SW_R; ++k) {
        st.h1[k] = 0u; st.h2[k] = 0u; st.sel[k] = 0u;
        st.prA[k] = prm.dummy; st.prB[k] = prm.dummy;
    }
    // Top boundary H[0][*] = 0 (source.cpp:44)
    constexpr int STRIPS = L / SW_R;
    static_assert(L >= 64 && (L & (L - 1)) == 0, "L must be a power of two >= 64");
    if (FAST && (V & SW_V_FIFO_PREOFF)) {
        for (int c = 0; c < L; ++c) fifo.push(c, prm.G * (uint32_t)(c + 1));   // a true zero in the frame of its pop: g(c+1)
    } else {
        for (int c = 0; c < L; ++c) fifo.push(c, 0u);
    }
    if (Fifo::kPrefetch) {
#pragma unroll
        for (int q = 0; q < (Fifo::kPrefetch == 2 ? 16 : 8); ++q) st.pq[q] = fifo.pop(q);
    }
    st.Z = FAST ? prm.G + prm.G : 0u;         // Z_0 = g*(0+2)
    st.Zp = FAST ? prm.G : 0u;                // Z_{-1}
    st.B = FAST ? prm.G : 0u;                 // best = 0 in the frame of step -1
    st.up0 = FAST ? prm.G : 0u;               // a true zero at step -1 (read once, as a diag of column 0, and overridden)
    st.dg0 = 0u;

    uint32_t an_lo[4], an_hi[4];
    ld8<(V & SW_V_COHERENT_LD) != 0>(b_lo, st.bq_lo[0], st.bq_lo[1]);
    ld8<(V & SW_V_COHERENT_LD) != 0>(b_lo + dqb, st.bq_hi[0], st.bq_hi[1]);
    ld16<(V & SW_V_COHERENT_LD) != 0>(a_lo, an_lo); ld16<(V & SW_V_COHERENT_LD) != 0>(a_lo + dqa, an_hi);

    for (int strip = 0; strip <= STRIPS; ++strip) {
        if (FAST && strip > 0) {
            // Renormalise the frame: T restarts at 0, every live value drops by L*g
            // (all are >= Z_{T-2} = L*g here, so the plain subtraction cannot borrow).
#pragma unroll
            for (int k = 0; k < SW_R; ++k) { st.h1[k] = fsub(st.h1[k], prm.C, prm); st.h2[k] = fsub(st.h2[k], prm.C, prm); }
            st.Z = fsub(st.Z, prm.C, prm); st.Zp = fsub(st.Zp, prm.C, prm);
            st.B = fsub(st.B, prm.C, prm); st.up0 = fsub(st.up0, prm.C, prm);
        }
        sw_iter16<FAST, true, L, V>(st, fifo, t4, prm, 0, b_lo, dqb, an_lo, an_hi, strip >= STRIPS);
        if (strip == STRIPS) break;
#pragma unroll 1
        for (int j = 1; j < L / 16; ++j) {
            if (j == L / 16 - 1) {   // one iteration ahead of the wrap that consumes them
                const int ns = (strip + 1 < STRIPS) ? strip + 1 : STRIPS - 1;
                ld16<(V & SW_V_COHERENT_LD) != 0>(a_lo + 16 * ns, an_lo); ld16<(V & SW_V_COHERENT_LD) != 0>(a_lo + 16 * ns + dqa, an_hi);
            }
            sw_iter16<FAST, false, L, V>(st, fifo, t4, prm, 16 * j, b_lo, dqb, an_lo, an_hi, false);
        }
    }
    if (FAST) {
        // B is in the frame of the last step; st.Zp is that step's Z.
        score_lo = (int32_t)(int16_t)(st.B & 0xffffu) - (int32_t)(int16_t)(st.Zp & 0xffffu);
        score_hi = (int32_t)(int16_t)(st.B >> 16) - (int32_t)(int16_t)(st.Zp >> 16);
    } else {
        score_lo = (int32_t)(int16_t)(st.B & 0xffffu);
        score_hi = (int32_t)(int16_t)(st.B >> 16);
    }
}

} // namespace swb
