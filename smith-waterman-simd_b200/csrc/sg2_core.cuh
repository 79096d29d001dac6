// sg2_core.cuh -- one round of the adaptive-banded X-drop semi-global aligner, FOUR LANES PER PAIR,
// eight band cells per lane in packed int16x2 registers.
//
// What it computes: exactly SemiGlobal_AdaptiveBanded_XDrop_111_32_70 (/root/reference/source.cpp:1836-1976)
// and hence its AVX2 forms (source.cpp:1978-2725): see sg_kernel.cuh for the statement of the algorithm.
//
// Why this shape.  A round is one serial chain (direction -> shift -> 32 cells -> maximum -> X-drop), so the
// bound is instructions per round.  With a warp per pair every instruction serves ONE cell per lane; here a
// lane owns cells 8q..8q+7 of the band as four int16x2 words, so the cell arithmetic, the shifts and the
// traceback masks are packed two cells per instruction and a warp advances EIGHT pairs per round.
//   * Values live in the X-drop frame: u = value - T with T = max(best - 70, 1) (the reference's own
//     trick for its 8-bit AVX2 forms, offset_diff at source.cpp:2661-2665), so every live cell is in 0..70
//     and int16 never overflows at any length.  A dropped cell is the sentinel F = 0x807F (-32641):
//         t2 = max(diag + sd, hor, ver)                   VIADD.16x2 + VIMNMX3.S16x2
//         R  = umin(max(t2 - c, F), F)                    VIADDMNMX.S16x2 + VIMNMX.U16x2
//     where sd = score + 1 - (T's last increment), c = 1 + (T's increment this round); the unsigned
//     minimum maps every negative value (unsigned >= 0x807F) to F and keeps 0..72 -- the X-drop and the
//     reference's "0 = dropped" in one instruction, with no compare/select.
//   * The band shift of the reference (alignr/permute2x128, source.cpp:2622,2632) is a funnel shift by
//     16 or 0 bits per word (the amount is the direction, so the pairs of a warp do not diverge) plus ONE
//     shuffle between neighbouring lanes that carries the boundary cell and the boundary base together.
//   * The 32 bases of either sequence under the band are two registers per lane (a byte per cell); they
//     shift by 8 or 0 bits; the match score of two cells is one PRMT through an 8-byte table indexed by
//     a XOR b (the reference's pshufb table, source.cpp:2640-2641).
//   * Traceback evidence, not band values, is stored (as in the warp-per-pair kernel): per cell "came from
//     the diagonal" (t2 == diag + sd) and "came from above" (t2 == ver), as two packed compares (HSET2)
//     gathered by one PRMT and three bit-selects: 4 bytes per lane per round, 16 bytes per pair per
//     round, pos_y in the spare bits.
//
// Every function is SWB_HD and templated on an Env that supplies the lane index and the two shuffles, so
// the same text runs on the device (real shuffles) and in tests/emu (four coroutines in lock step).
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif
#include "sw_core.cuh"

namespace swb {

#if defined(__CUDA_ARCH__)
// 0xFFFF in every half where a == b.  HSET2 compares the halves as fp16; that equals integer equality because the
// operands here are never NaN patterns (0x7C01.., 0xFC01..) and never 0x8000 (-0.0 == +0.0): live cells are
// 0..72 and everything else lies in 0x8070..0x8082 (F and the few values around it).
SWB_HD uint32_t veq2(uint32_t a, uint32_t b)
{
    return __heq2_mask(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
}
SWB_HD uint32_t vminu2(uint32_t a, uint32_t b) { return __vminu2(a, b); }        // VIMNMX.U16x2
SWB_HD uint32_t fsl(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }   // SHF.L.W
SWB_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }   // SHF.R.W
#else
SWB_HD uint32_t veq2(uint32_t a, uint32_t b) { return (((a ^ b) & 0xffffu) ? 0u : 0xffffu) | (((a ^ b) >> 16) ? 0u : 0xffff0000u); }
SWB_HD uint32_t vminu2(uint32_t a, uint32_t b)
{
    const uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
}
SWB_HD uint32_t fsl(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31u; return s ? (hi << s) | (lo >> (32u - s)) : hi; }
SWB_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31u; return s ? (lo >> s) | (hi << (32u - s)) : lo; }
#endif

constexpr int SG2_X = 70;                          // X_THRESHOLD, source.cpp:1848
constexpr uint32_t SG2_F = 0x807Fu;                // dropped / never reached
constexpr uint32_t SG2_FF = 0x807F807Fu;
constexpr uint32_t SG2_PAD_A = 0xC4u;              // base bytes: code * 0x11, seq1 side has bit 7 set; the pads (4, 5)
constexpr uint32_t SG2_PAD_B = 0x55u;              //   differ from every base and from each other (source.cpp:1913-1915)

struct Sg2State {
    uint32_t R[4];        // this round's cells 8q..8q+7 (word k = cells 2k | 2k+1 << 16), frame T, dropped = F
    uint32_t H[4], V[4];  // the previous round's left / upper neighbours (views of the round before it)
    uint32_t A[2], B[2];  // bases under the band: seq1 (byte c = cell c) and seq2
    uint32_t Rb[4];       // t2 of the best round (to find the end cell)
    uint32_t lut_lo, lut_hi;   // sd table: index 0 = match
    uint32_t right;       // the next round moves right (else down)
    uint32_t got;         // what enters this lane from its neighbour in the next round: bits 0-15 cell, 16-23 base
    uint32_t next_raw;    // lanes 0 and 3: the next base to enter the band (code, or 4 / 5 = pad), loaded a round ahead
    uint32_t role_base;   // F | (0x80 << 16 in lane 0)
    int32_t cidx;         // index of next_raw in seq1 (lane 0) / seq2 (lane 3)
    int32_t pos_y, pos_x; // the band's upper-right cell: (pos_y, pos_x - 31), source.cpp:1873-1874
    int32_t best, T, best_round, best_py, best_m;
};

SWB_HD int32_t sg2_half(uint32_t w, int hi) { return (int32_t)(int16_t)(hi ? (w >> 16) : (w & 0xffffu)); }

// Round 0 (source.cpp:1876-1884): cell 31 = X_THRESHOLD, everything else unreached; band at (0, 31).  Round 1
// always moves right (result[0] = 0 < result[31] = 70).
template <class Env>
SWB_HD void sg2_init(Sg2State& s, const Env& env, const uint8_t* seq1, const uint8_t* seq2, int len)
{
    const int q = env.q();
#pragma unroll
    for (int w = 0; w < 4; ++w) { s.R[w] = SG2_FF; s.H[w] = SG2_FF; s.V[w] = SG2_FF; }
    s.T = 1;                                         // max(70 - 70, 1)
    if (q == 3) s.R[3] = ((uint32_t)(SG2_X - 1) << 16) | SG2_F;
#pragma unroll
    for (int w = 0; w < 4; ++w) s.Rb[w] = s.R[w];
    // cell i holds seq1p[31 - i] = seq1[30 - i] (i = 31: pad) and seq2p[i] = pad
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        uint32_t a = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int idx = 30 - (8 * q + 4 * k + c);
            const uint32_t e = ((unsigned)idx < (unsigned)len) ? ((uint32_t)seq1[idx] * 0x11u | 0x80u) : SG2_PAD_A;
            a |= e << (8 * c);
        }
        s.A[k] = a;
        s.B[k] = SG2_PAD_B * 0x01010101u;
    }
    s.lut_lo = 2u; s.lut_hi = 0u;
    s.pos_y = 0; s.pos_x = 31;
    s.best = SG2_X; s.best_round = 0; s.best_py = 0; s.best_m = SG2_X - 1;
    // bases enter at cell 0 on a down move (lane 0: seq1p[pos_y + 31] = seq1[pos_y + 30]) and at cell 31 on a
    // right move (lane 3: seq2p[pos_x] = seq2[pos_x - 32]); round 1 takes seq2[0]
    s.role_base = SG2_F | (q == 0 ? 0x800000u : 0u);
    s.right = 1u;
    s.got = SG2_F | (SG2_PAD_B << 16);
    s.cidx = 31;
    s.next_raw = 4u;
    if (q == 0 && 31 < len) s.next_raw = seq1[31];
    if (q == 3) {
        s.got = SG2_F | ((0 < len ? (uint32_t)seq2[0] : 5u) * 0x110000u);
        s.cidx = 1;
        s.next_raw = (1 < len) ? (uint32_t)seq2[1] : 5u;
    }
}

// One round (source.cpp:1886-1942).  `role_seq` = seq1 in lane 0, seq2 in lane 3 (unused elsewhere);
// `rec_row` = this pair's records as uint32 [round][4].  Returns false when every cell is <= 0 (source.cpp:1938);
// a pair in that state is inert -- further rounds change neither its best nor its cells -- so the quads of a warp
// may keep running together until the last one is done.
// All communication of a round is ONE stage of seven independent shuffles issued right after the cells are
// computed (three for the maximum, two for the next direction, two for the boundary cell and base of either
// direction): everything that follows them is lane-local.  They carry t2, the cells BEFORE the X-drop; the
// receiver applies the drop itself, and the direction test is restated on t2:
//     result[0] < result[31]   <=>   t2[0] < t2[31]  and  t2[31] - c >= 0
// (with c the amount subtracted this round; dropped cells compare as the smallest value).
template <class Env>
SWB_HD bool sg2_round(Sg2State& s, Env& env, const uint8_t* role_seq, int len, int round, uint32_t* rec_row)
{
    const int q = env.q();
    const bool right = s.right != 0u;                      // source.cpp:1889
    const uint32_t got = s.got;
    uint32_t D[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) D[w] = right ? s.V[w] : s.H[w];       // source.cpp:1892,1903
    const uint32_t gh = got << 16, ga = got << 8, gb = got >> 16;
    const uint32_t sD = right ? 0u : 16u, sR = 16u - sD, cD = sD >> 1, cR = sR >> 1;
    s.H[3] = fsl(s.R[2], s.R[3], sD); s.H[2] = fsl(s.R[1], s.R[2], sD); s.H[1] = fsl(s.R[0], s.R[1], sD); s.H[0] = fsl(gh, s.R[0], sD);
    s.V[0] = fsr(s.R[0], s.R[1], sR); s.V[1] = fsr(s.R[1], s.R[2], sR); s.V[2] = fsr(s.R[2], s.R[3], sR); s.V[3] = fsr(s.R[3], got, sR);
    s.A[1] = fsl(s.A[0], s.A[1], cD); s.A[0] = fsl(ga, s.A[0], cD);
    s.B[0] = fsr(s.B[0], s.B[1], cR); s.B[1] = fsr(s.B[1], gb, cR);
    s.pos_y += right ? 0 : 1;
    s.pos_x += right ? 1 : 0;
    // ---- scores: selector byte of a cell = (a ^ b) in the low nibble, 8 | (a ^ b) in the high one (sign replication)
    const uint32_t x0 = s.A[0] ^ s.B[0], x1 = s.A[1] ^ s.B[1];
    uint32_t sd[4];
    sd[0] = prmt(s.lut_lo, s.lut_hi, x0); sd[1] = prmt(s.lut_lo, s.lut_hi, x0 >> 16);
    sd[2] = prmt(s.lut_lo, s.lut_hi, x1); sd[3] = prmt(s.lut_lo, s.lut_hi, x1 >> 16);
    // ---- the cells (source.cpp:1916-1926) and what the traceback will find for them (source.cpp:1960-1969)
    uint32_t t2[4], P[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t dsum = vadd2(D[w], sd[w]);
        t2[w] = vmax3(dsum, s.H[w], s.V[w]);
        P[w] = prmt(veq2(dsum, t2[w]), veq2(s.V[w], t2[w]), 0x6420u);       // 0xFF where equal: bytes d.lo d.hi u.lo u.hi
    }
    // ---- the shuffle stage
    uint32_t m = vmax2(vmax3(t2[0], t2[1], t2[2]), t2[3]);
    const uint32_t xr = prmt(t2[0], s.B[0], 0x0410u), xd = prmt(t2[3], s.A[1], 0x0732u);
    const uint32_t m1 = env.shfl_xor(m, 1), m2 = env.shfl_xor(m, 2), m3 = env.shfl_xor(m, 3);
    const uint32_t e0 = env.shfl(t2[0], 0), e31 = env.shfl(t2[3], 3);
    const uint32_t gn = env.shfl(xr, q + 1), gp = env.shfl(xd, q - 1);
    // ---- the record (independent of the shuffles)
    const uint32_t y01 = (P[0] & 0x55555555u) | (P[1] & 0xAAAAAAAAu), y23 = (P[2] & 0x55555555u) | (P[3] & 0xAAAAAAAAu);
    const uint32_t y = (y01 & 0x33333333u) | (y23 & 0xCCCCCCCCu);          // bit k of every nibble = word k
    const uint32_t spread = (((uint32_t)s.pos_y & 0xffu) << 4) | (((uint32_t)s.pos_y & 0xff00u) << 12);
    rec_row[4 * round + q] = (y & 0xF00FF00Fu) | spread;
    // ---- round maximum (source.cpp:1925)
    m = vmax2(vmax3(m, m1, m2), m3);
    const int32_t rmax = sg2_half(m, 0) > sg2_half(m, 1) ? sg2_half(m, 0) : sg2_half(m, 1);
    const int32_t amax = rmax - 1 + s.T;                   // with the reference's +70 offset
    if (amax > s.best) {                                   // strict: the FIRST round that reaches the best (source.cpp:1928-1931)
        s.best = amax; s.best_round = round; s.best_py = s.pos_y; s.best_m = rmax;
#pragma unroll
        for (int w = 0; w < 4; ++w) s.Rb[w] = t2[w];
    }
    const int32_t Tn = (s.best - SG2_X > 1) ? s.best - SG2_X : 1;
    const uint32_t off = (uint32_t)(Tn - s.T);             // 0 or 1
    s.T = Tn;
    // ---- X-drop and "<= 0 is dropped" (source.cpp:1918,1933-1936) in the new frame
    const uint32_t nc = 0xFFFFFFFFu - off * 0x00010001u;   // (-c, -c), c = 1 + off
#pragma unroll
    for (int w = 0; w < 4; ++w) s.R[w] = vminu2(vaddmax2(t2[w], nc, SG2_FF), SG2_FF);
    // ---- the next round's direction and what enters this lane then
    const int32_t t31 = sg2_half(e31, 1);
    const bool rn = sg2_half(e0, 0) < t31 && t31 > (int32_t)off;          // t31 - c >= 0
    const bool edge = rn ? (q == 3) : (q == 0);            // the band's end: a dropped cell and a new base come in
    uint32_t cand = rn ? gn : gp;
    if (edge) {
        cand = s.next_raw * 0x110000u + s.role_base;
        ++s.cidx;
        s.next_raw = (q == 0) ? 4u : 5u;
        if ((uint32_t)s.cidx < (uint32_t)len) s.next_raw = role_seq[s.cidx];
    }
    s.got = vminu2(vaddmax2(cand, nc & 0xffffu, SG2_FF), SG2_FF);          // drop the cell, leave the base
    s.right = rn ? 1u : 0u;
    // next round's diagonal inputs are one frame older: sd = score + 1 - off
    const uint32_t noff = 0u - off;
    s.lut_lo = (noff & 0xFFFFFF03u) ^ 2u;                  // off = 0: 02 00 00 00, off = 1: 01 FF FF FF
    s.lut_hi = noff;
    return amax > 0;
}

// After the last round: the end cell is the upper-right-most cell of the best round that holds the best
// score (source.cpp:1953-1954).  Returns this lane's word of record 0 = {best round, its pos_y, end_y, end_x}.
template <class Env>
SWB_HD uint32_t sg2_finish(const Sg2State& s, Env& env, int32_t& score, int32_t& end_y, int32_t& end_x)
{
    const int q = env.q();
    int32_t loc = -1;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (sg2_half(s.Rb[c >> 1], c & 1) == s.best_m) loc = 8 * q + c;
    int32_t o = (int32_t)env.shfl_xor((uint32_t)loc, 1); loc = loc > o ? loc : o;
    o = (int32_t)env.shfl_xor((uint32_t)loc, 2); loc = loc > o ? loc : o;
    score = s.best - SG2_X;
    end_y = s.best_py + 31 - loc;
    end_x = (s.best_round - s.best_py) - 31 + loc;         // pos_x = 31 + (number of right moves)
    return (uint32_t)(q == 0 ? s.best_round : q == 1 ? s.best_py : q == 2 ? end_y : end_x);
}

// One traceback step on a record (source.cpp:1958-1971): diagonal first, then up, else left.
// rec = the four lane words of round r = y + x.  Returns the op: 0 = diagonal, 1 = down (y+1), 2 = right (x+1).
SWB_HD uint32_t sg2_tb_step(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, int& y, int& x, int& r)
{
    const int py = (int)(((r0 >> 4) & 0xffu) | ((r0 >> 12) & 0xff00u));
    const int o = 31 - (y - py);                           // band element of (y, x) in round r (source.cpp:1947)
    const uint32_t w = (o & 16) ? ((o & 8) ? r3 : r2) : ((o & 8) ? r1 : r0);
    const int pos = ((o & 7) >> 1) + 12 * (o & 1);
    const uint32_t d = (w >> pos) & 1u;
    const uint32_t u = (w >> (pos + 16)) & 1u & ~d;
    y -= (int)(d | u);
    x -= (int)(1u - u);
    r -= 1 + (int)d;
    return 2u - 2u * d - u;
}

} // namespace swb
